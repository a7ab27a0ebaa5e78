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
                 precision='bf16'):
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

    def update(self, batch):
        raise NotImplementedError('SAC updates are outside the hot path of this package')
