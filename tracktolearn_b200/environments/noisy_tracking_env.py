"""NoisyTrackingEnvironment (reference: environments/noisy_tracking_env.py:9-77).

Every ``ttl_track`` run uses this class, even with ``--noise 0``: the float64
``rng.normal`` array added to the float32 actions makes the whole direction arithmetic
float64, which the step kernel reproduces with ``dir_f64 = 1``.
"""
import numpy as np
import torch

from tracktolearn_b200.environments.tracking_env import TrackingEnvironment


class NoisyTrackingEnvironment(TrackingEnvironment):

    def __init__(self, dataset_file, split_id, env_dto):
        self.noise = env_dto['noise']
        self.fa_map = None
        if env_dto.get('fa_map'):
            # The reference parses --fa_map but never reaches this branch (ttl_track.py:86 looks
            # for the wrong key); FA-scaled noise is dead code there and absent here.
            raise NotImplementedError('FA-scaled noise is dead code in the reference (F13)')
        self.max_action = 1.
        self._float64_directions = True
        super().__init__(dataset_file, split_id, env_dto)

    def load_subject(self):
        super().load_subject()
        self._params.dir_f64 = 1

    def _draw_noise(self, shape):
        """noisy_tracking_env.py:74: host RandomState draw, float64; None when sigma == 0
        (adding an all-zero float64 array only changes the dtype, which dir_f64 covers)."""
        if self.noise > 0.:
            noise = self.rng.normal(0., self.noise, size=shape)
            return torch.from_numpy(noise).to(self.device)
        return None

    def step(self, actions):
        """Reference: noisy_tracking_env.py:38-77."""
        actions = self._host_actions(actions)
        return self._step(actions, self._draw_noise((self._n_alive_host, 3)))

    def step_device(self, actions, noise=None):
        if noise is None and self.noise > 0.:
            noise = self._draw_noise((self._n_alive_host, 3))
        return super().step_device(actions, noise)
