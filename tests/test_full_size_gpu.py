"""BASELINE.json configs[1] (145x174x145, 1.25 mm) and configs[2] (290^3, 0.5 mm: a 4.7 GB channel-padded
volume, byte offsets beyond 2^32) at full size with 50 000 slots: size-independent properties of a whole
tracked batch, and the stopping criteria / state of a sample of its streamlines re-evaluated by the CPU
oracle on the device's own points ("identical inputs": the oracle sees exactly the streamlines the device
produced, so no closed-loop drift enters the comparison).  The actor runs in the product's default tier
(fp16 tensor cores)."""
import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = {'configs1_145x174x145': ((145, 174, 145), 1.25), 'configs2_290cubed': ((290, 290, 290), 0.5)}
N_SEEDS = 120000
N_ACTOR = 50000


@pytest.fixture(scope='module', params=list(CONFIGS))
def tracked(request):
    SHAPE, VOXEL_MM = CONFIGS[request.param]
    STEP_MM = VOXEL_MM / 0.9987237 * 0.75
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import NoisyTrackingEnvironment
    from tracktolearn_b200.environments.utils import random_seeds_from_mask
    from tracktolearn_b200.tracking.tracker import Tracker
    dev = torch.device('cuda:0')
    sub = synthetic.make_subject(SHAPE, seed=1234, device=dev, with_peaks=False)
    affine = np.diag([VOXEL_MM] * 3 + [1.0])
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['seed_mask'], affine), None, affine)
    dto = {'n_dirs': 100, 'theta': 30.0, 'npv': 1, 'binary_stopping_threshold': 0.1, 'step_size': STEP_MM,
           'min_length': 10.0, 'max_length': 300.0, 'oracle_checkpoint': None,
           'oracle_stopping_criterion': False, 'scoring_data': None, 'compute_reward': False,
           'alignment_weighting': 0.0, 'oracle_bonus': 0.0, 'rng': np.random.RandomState(1337), 'device': dev,
           'target_sh_order': 8, 'noise': 0.0, 'fa_map': None, 'state_of_stopped': False}
    env = NoisyTrackingEnvironment(subject, 'testing', dto)
    rs = np.random.RandomState(4242)
    seeds = random_seeds_from_mask(sub['seed_mask'].cpu().numpy(), 3, rs)
    rs.shuffle(seeds)
    env.seeds = seeds[:N_SEEDS]
    alg = SACAuto(615, 3, '1024-1024-1024', n_actors=N_ACTOR, device=dev, precision='fp16')
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(615, '1024-1024-1024', seed=1111, kind='tracking'))
    tracker = Tracker(alg, N_ACTOR, min_length=10.0, max_length=300.0)
    batches = list(tracker.track_packed(env, copy=True))
    assert len(batches) == 1
    host = {'sh': sub['sh'].cpu().numpy(), 'mask': sub['mask'].cpu().numpy(), 'step_vox': STEP_MM / VOXEL_MM}
    del sub
    yield env, batches[0], host
    env._batch = None
    torch.cuda.empty_cache()


def test_whole_batch_properties(tracked):
    env, t, _ = tracked
    n = len(env.seeds)
    assert len(t.lengths) == n == N_SEEDS
    lens = np.asarray(t.lengths)
    flags = np.asarray(t.data_per_streamline['flags'])
    assert flags.min() > 0                                   # every streamline ended on a criterion
    assert set(np.unique(flags)) <= {1, 2, 4, 5, 3, 6, 7}
    raw = env.lengths                                        # points before the last-point trim
    assert raw.min() >= 2 and raw.max() <= env.max_nb_steps
    trim = (flags & 5) != 0                                  # CURVATURE | MASK drop their last point
    np.testing.assert_array_equal(lens, raw - trim)
    assert ((flags & 2) != 0).sum() == (raw == env.max_nb_steps).sum()
    assert env.streamline_steps() == int((raw - 1).sum())    # device counter == work done
    # every streamline starts on its seed (float32 of the float64 seed)
    first = t.data[t.offsets[:-1]]
    np.testing.assert_array_equal(first, env.seeds.astype(np.float32))
    # every segment has the step length r.  Positions are float32 like the reference's streamline buffer
    # (tracking_env.py:116): at coordinates up to 290 one ulp is 3e-5 voxel, so a segment length recomputed
    # from two stored points is off by up to sqrt(3) ulp; 1e-5 (north_star's position tolerance) where the
    # coordinates allow it.
    seg = np.linalg.norm(np.diff(t.data.astype(np.float64), axis=0), axis=1)
    inner = np.ones(len(seg), dtype=bool)
    inner[t.offsets[1:-1] - 1] = False
    r = tracked[2]['step_vox']
    tol = max(1e-5, float(np.sqrt(3.0) * np.spacing(np.float32(t.data.max()))))
    assert np.abs(seg[inner] - r).max() < tol
    assert env.n_alive() == 0
    assert lens.mean() > 20                                  # the synthetic field is trackable


def test_sampled_streamlines_against_oracle_criteria_and_state(tracked):
    env, t, sub = tracked
    rs = np.random.RandomState(1)
    pick = rs.choice(len(t.lengths), size=400, replace=False)
    crit = O.BinaryStoppingCriterion(sub['mask'].astype(np.uint8), 0.1)
    theta, max_nb = 30.0, env.max_nb_steps
    pts_dev = env._batch.points
    flags = np.asarray(t.data_per_streamline['flags'])
    raw = env.lengths
    states_in, states_L = [], []
    for i in pick:
        L = int(raw[i])
        full = pts_dev[i, :L].cpu().numpy()[None]            # with the point that triggered the stop
        f = 0
        if O.is_too_long(full, max_nb)[0]:
            f |= O.LENGTH
        if O.is_too_curvy(full, theta)[0]:
            f |= O.CURVATURE
        if crit(full)[0]:
            f |= O.MASK
        assert f == flags[i], (i, L, f, flags[i])
        if L > 2:                                            # one step earlier it was still alive
            prev = full[:, :L - 1]
            assert not O.is_too_long(prev, max_nb)[0]
            assert not O.is_too_curvy(prev, theta)[0]
            assert not crit(prev)[0]
        if len(states_in) < 64 and L >= 5:
            states_in.append(full[0, :5])
            states_L.append(5)
    # _format_state on the full-size volume for 64 of them (fp32 tier, 1e-5 absolute)
    from tracktolearn_b200 import _lib
    import ctypes
    P = np.stack(states_in).astype(np.float32)
    want = O.format_state(sub['sh'], P, O.neighborhood_directions(env.step_size), 100)
    got = env._format_state(P) if hasattr(env, '_format_state') else None
    assert got is not None
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
    assert _lib is not None and ctypes is not None
