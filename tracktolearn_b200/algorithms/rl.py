"""Episode drivers (reference: algorithms/rl.py:58-106, the act -> step -> harvest loop).

The reference crosses the host/device boundary three times per step (action D2H with a
sync, coordinates H2D, previous directions H2D).  Here the loop only enqueues kernels: the
actor reads the alive count from device memory, the env step consumes the action tensor in
place, and the host looks at the alive count every ``sync_every`` steps.
"""
import numpy as np
import torch


class RLAlgorithm(object):
    """Reference: algorithms/rl.py:8-56 (constructor arguments kept)."""

    def __init__(self, input_size, action_size=3, hidden_size=256, lr=3e-4, gamma=0.99,
                 batch_size=10000, rng=None, device=None):
        self.max_action = 1.
        self.t = 1
        self.action_size = action_size
        self.lr = lr
        self.gamma = gamma
        self.device = device
        self.batch_size = batch_size
        self.rng = rng
        self.sync_every = 8

    def validation_episode(self, initial_state, env, prob=1., max_steps=None, on_step=None):
        """Run the agent until every streamline of the env's current batch is done.

        Reference: algorithms/rl.py:58-106.  ``initial_state`` is accepted for signature
        compatibility; the state rows are read from the env's device buffers.  Returns the
        cumulative reward (0 when the env does not compute rewards)."""
        actor = self.agent.actor
        running_reward = 0
        reward_acc = None
        n_up = env._n_alive_host
        if n_up == 0:
            return running_reward
        action_buf = torch.empty((env._b.n_slots, self.action_size), dtype=torch.float32,
                                 device=env.device)
        it = 0
        limit = max_steps if max_steps is not None else 1 << 30
        while n_up > 0 and it < limit:
            state = env.current_state()          # None when the env produces bf16 rows only
            rows = n_up
            actor.forward_device(state, prob, n_rows_dev=env.alive_count_tensor(), n_rows=rows,
                                 want_logp=False, out_action=action_buf,
                                 state_bf16=env.current_state_bf16(), layout=env.bf16_layout)
            env.step_device(action_buf)
            if env.compute_reward:
                r = env._batch.reward[:rows].sum(dtype=torch.float64)
                reward_acc = r if reward_acc is None else reward_acc + r
            env.harvest_device()
            it += 1
            if on_step is not None:
                on_step(it)
            if it % self.sync_every == 0:
                env.n_alive()
                n_up = env._n_alive_host
        env.n_alive()
        if reward_acc is not None:
            running_reward = float(reward_acc.item())
        self.last_episode_steps = it
        return running_reward
