"""Host helpers around the environment (seeding)."""
import numpy as np


def random_seeds_from_mask(mask, npv, rng=None):
    """dipy ``random_seeds_from_mask(mask, np.eye(4), seeds_count=npv)`` as the reference calls
    it (environments/env.py:216-219): for i in 1..npv, for every mask voxel in C order,
    ``voxel + random(3) - 0.5`` from the global numpy RNG.  Vectorised: one draw of shape
    [npv, n_voxels, 3] consumes the stream in the same order.  float64, voxel centres at
    integer coordinates."""
    where = np.argwhere(np.asarray(mask, dtype=bool))
    draw = (rng.random_sample if rng is not None else np.random.random_sample)
    grid = draw((npv, len(where), 3))
    seeds = where[None, :, :].astype(np.float64) + grid - 0.5
    return seeds.reshape(-1, 3)
