"""Builds product envs from golden fixtures (GPU tests only)."""
import numpy as np
import torch

from tests.helpers import meta, subject_for
from tracktolearn_b200.datasets.utils import MRIDataVolume


def make_gpu_env(g, noisy, compute_reward, sub=None, seeds=None, oracle_checkpoint=None,
                 oracle_stopping=False, oracle_bonus=0.0, min_length=1.0, oracle_precision='fp32', **over):
    from tracktolearn_b200.environments import NoisyTrackingEnvironment, TrackingEnvironment
    sub = sub or subject_for(g)
    m = meta(g)
    m.update(over)
    affine = np.diag([m['vox']] * 3 + [1.0])
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['mask'], affine), MRIDataVolume(sub['peaks'], affine), affine)
    dto = {'n_dirs': 100, 'theta': m['theta'], 'npv': 1, 'binary_stopping_threshold': m['threshold'],
           'step_size': m['step_mm'], 'min_length': min_length, 'max_length': m['max_length'],
           'oracle_checkpoint': oracle_checkpoint, 'oracle_stopping_criterion': oracle_stopping,
           'scoring_data': None,
           'compute_reward': compute_reward, 'alignment_weighting': 1.0, 'oracle_bonus': oracle_bonus,
           'rng': np.random.RandomState(1337), 'device': torch.device('cuda:0'), 'target_sh_order': 8,
           'noise': 0.0, 'fa_map': None, 'oracle_precision': oracle_precision}
    cls = NoisyTrackingEnvironment if noisy else TrackingEnvironment
    env = cls(subject, 'testing', dto)
    if seeds is not None:
        env.seeds = np.asarray(seeds)
    return env, sub
