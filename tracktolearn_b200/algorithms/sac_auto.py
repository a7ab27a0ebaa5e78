"""SACAuto agent holder for tracking (reference: algorithms/sac_auto.py:37-137).

Only what the tracking path needs: the constructor signature ``ttl_track.py`` uses
(:159-165) and ``.agent`` (a ``SACActorCritic``).  The reference also allocates a 1e6-row
pinned replay buffer here that tracking never touches (sac_auto.py:134); we do not.
Training (``update``) is outside this package's hot path -- see DESIGN.md.
"""
import numpy as np
import torch

from tracktolearn_b200.algorithms.rl import RLAlgorithm
from tracktolearn_b200.algorithms.shared.offpolicy import SACActorCritic


class SACAuto(RLAlgorithm):

    def __init__(self, input_size, action_size, hidden_dims, lr=3e-4, gamma=0.99, alpha=0.2,
                 n_actors=4096, batch_size=2 ** 12, replay_size=1e6, rng=None, device=None,
                 precision='fp16'):
        super().__init__(input_size, action_size, hidden_dims, lr, gamma, batch_size, rng, device)
        self.n_actors = n_actors
        self.agent = SACActorCritic(input_size, action_size, hidden_dims, device, precision=precision)
        self.alpha = alpha
        self.start_timesteps = 80000
        self.total_it = 0
        self.tau = 0.005
        self.replay_size = replay_size

    def sample_action(self, state):
        """Reference: algorithms/sac.py:123-133."""
        return self.agent.select_action(state, probabilistic=1.0)

    # ------------------------------------------------------------------------------ training
    def enable_training(self, lr=3e-4, gamma=0.99, replay_size=None, batch_size=None, start_timesteps=None):
        """Build the learner (torch autograd networks + Adam, sac_train.py), make the tensor-core
        actor read ITS weights, and allocate the device-resident replay buffer."""
        from tracktolearn_b200.algorithms.sac_train import SACAutoLearner
        from tracktolearn_b200.algorithms.shared.replay import OffPolicyReplayBuffer
        actor = self.agent.actor
        hidden = '-'.join(str(int(w)) for w in actor.hidden_layers)
        self.learner = SACAutoLearner(actor.state_dim, actor.action_dim, hidden, lr=lr, gamma=gamma,
                                      alpha=self.alpha, device=actor.device)
        self.learner.actor.load_state_dict(actor.state_dict())
        self.learner.target_actor.load_state_dict(actor.state_dict())
        self.agent.attach_learner(self.learner)        # a critic loaded from a checkpoint moves into the learner
        self.learner.broadcast_parameters()
        actor.share_parameters(self.learner.actor.state_dict())
        if replay_size is not None:
            self.replay_size = replay_size
        if batch_size is not None:
            self.batch_size = batch_size
        if start_timesteps is not None:
            self.start_timesteps = start_timesteps
        self.replay_buffer = OffPolicyReplayBuffer(actor.state_dim, actor.action_dim, max_size=int(self.replay_size),
                                                   device=self.device)
        return self.learner

    def update(self, batch):
        """Reference: sac_auto.py:139-250."""
        losses = self.learner.update(batch)
        self.agent.actor.refresh_weights()
        self.total_it += 1
        return losses

    def _episode(self, initial_state, env):
        """Rollout + update loop (reference: algorithms/ddpg.py:141-232): sample an action for every
        alive streamline, step the env, push the transitions of all of them to the replay buffer,
        do ONE gradient update per environment step once ``start_timesteps`` transitions have been
        seen, harvest.  Everything stays on the device; the only host read per step is the alive
        count."""
        import ctypes
        from tracktolearn_b200 import _lib
        actor = self.agent.actor
        running_reward = 0.0
        episode_length = 0
        losses_log = []
        A = env._n_alive_host
        S = env.get_state_size()
        import torch.distributed as dist
        dp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        while True:
            if A > 0:
                state = env.current_state()[:A]
                action, _, _ = actor.forward_device(state, 1.0, n_rows=A, want_logp=False)
                env.step_device(action)
                next_state = torch.empty((A, S), dtype=torch.float32, device=env.device)
                _lib.check(env._lib.ttl_env_gather_step_state(ctypes.byref(env._b), env._cur, A, _lib.ptr(next_state),
                                                              S, _lib.stream_ptr(env.device)), 'ttl_env_gather_step_state')
                reward = env._batch.reward[:A] if env.compute_reward else torch.zeros((A,), device=env.device)
                done = env._batch.stop[:A]
                self.replay_buffer.add(state, action, next_state, reward, done)
                running_reward += float(reward.sum().item()) if env.compute_reward else 0.0
            ready = self.t >= self.start_timesteps
            if dp:
                # replicas must issue the same sequence of gradient all-reduces: update only when every
                # rank has enough transitions, and keep updating (from the local replay shard) until
                # the longest episode among the ranks has ended
                flag = torch.tensor([float(A), -float(ready)], device=env.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                any_alive, ready = float(flag[0]) > 0, float(flag[1]) < 0
            else:
                any_alive = A > 0
            if ready and (A > 0 or dp) and any_alive:
                losses_log.append(self.update(self.replay_buffer.sample(self.batch_size)))
            if A > 0:
                self.t += A
                env.harvest_device()
                A = env.n_alive()
                episode_length += 1
            if not any_alive:
                break
        return running_reward, losses_log, episode_length, {}
