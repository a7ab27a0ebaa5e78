"""CUDA tracking step (through the C ABI) against the reference-recorded fixtures and the
CPU oracle.  Tolerances (north_star): positions 1e-5, state 1e-5 (fp32), flags/dones/lengths
bit-exact."""
import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O
from tests.helpers import load_golden, meta, split_by_counts, subject_for

pytestmark = pytest.mark.gpu
STATE_TOL = 1e-5
POS_TOL = 1e-5


@pytest.mark.parametrize('name,noisy,reward', [('env_noisy', True, False),
                                               ('env_plain_reward', False, True)])
def test_episode_matches_reference_fixture(name, noisy, reward):
    from tests.gpu_helpers import make_gpu_env
    g = load_golden(name)
    env, _ = make_gpu_env(g, noisy, reward, seeds=g['seeds'])
    assert env.max_nb_steps == meta(g)['max_nb_steps']
    n = len(g['seeds'])
    state = env.reset(0, n)
    np.testing.assert_allclose(state.cpu().numpy(), g['state_reset'], rtol=0, atol=STATE_TOL)
    counts = g['alive_counts']
    ci_g = split_by_counts(g['continue_idx'], counts)
    dones_g = split_by_counts(g['dones'], counts)
    pts_g = split_by_counts(g['new_points'], counts)
    flags_g = split_by_counts(g['step_flags'], counts)
    rew_g = split_by_counts(g['rewards'], counts)
    for t in range(int(g['n_steps'])):
        ci = env.continue_idx.copy()
        np.testing.assert_array_equal(ci, ci_g[t])
        st, r, done, info = env.step(g['actions'][t][ci])
        np.testing.assert_array_equal(done.astype(np.uint8), dones_g[t])
        np.testing.assert_array_equal(info['continue_idx'], ci_g[t])
        pts = env.streamlines[ci, env.length - 1]
        np.testing.assert_allclose(pts, pts_g[t], rtol=0, atol=POS_TOL, equal_nan=True)
        np.testing.assert_array_equal(env.flags[ci], flags_g[t])
        if reward:
            np.testing.assert_allclose(r, rew_g[t], rtol=0, atol=2e-6)
        else:
            assert r.shape == (n,) and not r.any()
        if 'state_%d' % t in g:
            np.testing.assert_allclose(st.cpu().numpy(), g['state_%d' % t], rtol=0, atol=STATE_TOL,
                                       equal_nan=True)
        hs, not_stopping = env.harvest()
        np.testing.assert_array_equal(not_stopping, ~done)
        if 'harvest_state_%d' % t in g:
            np.testing.assert_allclose(hs.cpu().numpy(), g['harvest_state_%d' % t], rtol=0,
                                       atol=STATE_TOL, equal_nan=True)
    assert len(env.continue_idx) == 0
    np.testing.assert_array_equal(env.flags, g['final_flags'])
    np.testing.assert_array_equal(env.lengths, g['final_lengths'])
    tr = env.get_streamlines()
    np.testing.assert_array_equal(tr.lengths, g['sl_lengths'])
    np.testing.assert_allclose(tr.data, g['sl_points'], rtol=0, atol=POS_TOL, equal_nan=True)
    np.testing.assert_array_equal(tr.data_per_streamline['flags'], g['final_flags'])


def test_edges_state_flags_reward():
    from tests.gpu_helpers import make_gpu_env
    g = load_golden('edges')
    env, sub = make_gpu_env(g, False, True)
    pts = g['points']
    L = pts.shape[1]
    for Lk in (1, 2, 3, L):
        sp = np.ascontiguousarray(pts[:, -Lk:])
        st = env._format_state(sp).cpu().numpy()
        np.testing.assert_allclose(st, g['state_L%d' % Lk], rtol=0, atol=STATE_TOL, equal_nan=True)
        stop, flags, mval, rew = env._compute_stopping_flags(sp, with_reward=True)
        np.testing.assert_array_equal(flags, g['flags_L%d' % Lk])
        np.testing.assert_array_equal(stop.astype(np.uint8), g['stop_L%d' % Lk])
        np.testing.assert_allclose(rew.astype(np.float64), g['reward_L%d' % Lk], rtol=0, atol=2e-6)
        if Lk == L:
            np.testing.assert_allclose(mval, g['mask_values'], rtol=0, atol=1e-13)


@pytest.mark.parametrize('noisy', [True, False])
def test_long_episode_matches_oracle(noisy):
    """Bigger batch, oracle-driven comparison: 40x44x36 volume, 1500 seeds, random-walk actions, every
    step compared (points, flags, dones, alive order).  (The exact BASELINE configs[0] shape -- 64^3,
    4096 streamlines -- is tests/test_tracker_gpu.py::test_config0_shape_closed_loop_matches_oracle.)"""
    from tests.gpu_helpers import make_gpu_env
    from tracktolearn_b200 import synthetic
    shape = (40, 44, 36)
    sub = {k: (v.numpy() if v is not None else None) for k, v in synthetic.make_subject(shape, seed=77).items()}
    rs = np.random.RandomState(5)
    seeds = synthetic.seeds_from_mask(synthetic.ellipsoid_mask(shape, frac=0.33).numpy(), 1, rs)
    rs.shuffle(seeds)
    seeds = seeds[:1500]
    n = len(seeds)
    g = {'meta_shape': np.asarray(shape), 'meta': np.asarray([1.0, 0.75, 30.0, 24.0, 0.1, 32.0, 0.75])}
    env, _ = make_gpu_env(g, noisy, True, sub=sub, seeds=seeds)
    ref = O.OracleEnv(sub['sh'], sub['mask'], seeds, 1.0, 0.75, theta=30.0, max_length_mm=24.0,
                      peaks=sub['peaks'], compute_reward=True, noisy=noisy)
    s_gpu = env.reset(0, n)
    s_ref = ref.reset(0, n)
    np.testing.assert_allclose(s_gpu.cpu().numpy(), s_ref, rtol=0, atol=STATE_TOL)
    a = rs.normal(size=(n, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = 0
    mismatched_flags = 0
    while len(ref.continue_idx):
        ci = ref.continue_idx.copy()
        np.testing.assert_array_equal(env.continue_idx, ci)
        a = a + 0.17 * rs.normal(size=(n, 3))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        act = (a[ci] * rs.uniform(0.3, 1.0, size=(len(ci), 1))).astype(np.float32)
        st_g, r_g, d_g, _ = env.step(act)
        st_r, r_r, d_r, _ = ref.step(act)
        np.testing.assert_allclose(env.streamlines[ci, env.length - 1], ref.streamlines[ci, ref.length - 1],
                                   rtol=0, atol=POS_TOL)
        np.testing.assert_array_equal(d_g, d_r)
        np.testing.assert_array_equal(env.flags[ci], ref.flags[ci])
        np.testing.assert_allclose(r_g, r_r, rtol=0, atol=2e-6)
        np.testing.assert_allclose(st_g.cpu().numpy(), st_r, rtol=0, atol=STATE_TOL)
        h_g, _ = env.harvest()
        h_r, _ = ref.harvest()
        np.testing.assert_allclose(h_g.cpu().numpy(), h_r, rtol=0, atol=STATE_TOL)
        t += 1
    assert t >= 25
    np.testing.assert_array_equal(env.lengths, ref.lengths)
    tr = env.get_streamlines()
    sl, _, fl = ref.get_streamlines()
    np.testing.assert_array_equal(tr.lengths, [len(s) for s in sl])
    np.testing.assert_allclose(tr.data, np.concatenate(sl), rtol=0, atol=POS_TOL)
    fl = np.asarray(fl)
    assert (fl & 1).any() and (fl & 2).any() and (fl & 4).any(), 'all three criteria should fire'


def test_device_protocol_equals_reference_protocol():
    """step_device/harvest_device (no host traffic) must leave the same device state as
    step/harvest; also covers skipping the state rows of stopped streamlines."""
    from tests.gpu_helpers import make_gpu_env
    g = load_golden('env_noisy')
    envA, sub = make_gpu_env(g, True, False, seeds=g['seeds'])
    envB, _ = make_gpu_env(g, True, False, sub=sub, seeds=g['seeds'])
    envB.state_of_stopped = False
    envB.load_subject()
    envB.seeds = g['seeds']
    n = len(g['seeds'])
    envA.reset(0, n)
    envB.reset(0, n)
    for t in range(int(g['n_steps'])):
        ci = envA.continue_idx.copy()
        envA.step(g['actions'][t][ci])
        hs, _ = envA.harvest()
        act = torch.from_numpy(g['actions'][t][ci]).cuda()
        envB.step_device(act)
        envB.harvest_device()
        nb = envB.n_alive()
        assert nb == hs.shape[0]
        np.testing.assert_array_equal(envB.current_state().cpu().numpy(), hs.cpu().numpy())
    np.testing.assert_array_equal(envA.flags, envB.flags)
    np.testing.assert_array_equal(envA.lengths, envB.lengths)
    np.testing.assert_array_equal(envA.get_streamlines().data, envB.get_streamlines().data)


def test_format_state_generic_channel_count():
    """Order-6 volume (28 coefficients): exercises the non-specialised state path."""
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import TrackingEnvironment
    rs = np.random.RandomState(4)
    shape = (10, 11, 12)
    sh = rs.normal(size=shape + (28,)).astype(np.float32)
    mask = np.ones(shape, dtype=np.uint8)
    affine = np.eye(4)
    dto = {'n_dirs': 100, 'theta': 30.0, 'npv': 1, 'binary_stopping_threshold': 0.1, 'step_size': 0.6,
           'min_length': 1.0, 'max_length': 12.0, 'oracle_checkpoint': None,
           'oracle_stopping_criterion': False, 'scoring_data': None, 'compute_reward': False,
           'alignment_weighting': 0.0, 'oracle_bonus': 0.0, 'rng': np.random.RandomState(0),
           'device': torch.device('cuda:0'), 'target_sh_order': 6, 'noise': 0.0, 'fa_map': None}
    env = TrackingEnvironment((MRIDataVolume(sh, affine), MRIDataVolume(mask, affine),
                               MRIDataVolume(mask, affine), None, affine), 'testing', dto)
    assert env.get_state_size() == 7 * 28 + 300
    pts = (rs.uniform(-1.0, 12.0, size=(200, 1, 3)) + np.cumsum(rs.normal(scale=0.4, size=(200, 9, 3)), 1)).astype(np.float32)
    ref = O.format_state(sh, pts, O.neighborhood_directions(0.6), 100)
    got = env._format_state(pts).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=STATE_TOL)
    # and through reset/step with the oracle env
    seeds = rs.uniform(2, 8, size=(64, 3))
    env.seeds = seeds
    renv = O.OracleEnv(sh, mask, seeds, 1.0, 0.6, max_length_mm=12.0, noisy=False)
    np.testing.assert_allclose(env.reset(0, 64).cpu().numpy(), renv.reset(0, 64), rtol=0, atol=STATE_TOL)
    act = rs.normal(size=(64, 3)).astype(np.float32)
    sg, _, dg, _ = env.step(act)
    sr, _, dr, _ = renv.step(act)
    np.testing.assert_array_equal(dg, dr)
    np.testing.assert_allclose(sg.cpu().numpy(), sr, rtol=0, atol=STATE_TOL)


def test_episode_with_oracle_criterion_and_bonus_matches_reference_fixture():
    """S9 / S13: TractOracle-Net consulted inside step() (stopping criterion from 5*min_nb_steps
    points on, sparse bonus for finished streamlines the oracle likes)."""
    from tests.gpu_helpers import make_gpu_env
    from tests.helpers import oracle_ckpt_for
    from tracktolearn_b200.oracles.oracle import OracleSingleton
    g = load_golden('env_oracle')
    OracleSingleton.clear()
    env, _ = make_gpu_env(g, False, True, seeds=g['seeds'], oracle_checkpoint=oracle_ckpt_for(g),
                          oracle_stopping=True, oracle_bonus=10.0, min_length=1.6)
    assert env.min_nb_steps == int(g['min_nb_steps'])
    n = len(g['seeds'])
    env.reset(0, n)
    counts = g['alive_counts']
    ci_g = split_by_counts(g['continue_idx'], counts)
    dones_g = split_by_counts(g['dones'], counts)
    flags_g = split_by_counts(g['step_flags'], counts)
    rew_g = split_by_counts(g['rewards'], counts)
    for t in range(int(g['n_steps'])):
        ci = env.continue_idx.copy()
        np.testing.assert_array_equal(ci, ci_g[t])
        st, r, done, _ = env.step(g['actions'][t][ci])
        np.testing.assert_array_equal(done.astype(np.uint8), dones_g[t])
        np.testing.assert_array_equal(env.flags[ci], flags_g[t])
        np.testing.assert_allclose(r, rew_g[t], rtol=0, atol=1e-5)
        if 'state_%d' % t in g:
            np.testing.assert_allclose(st.cpu().numpy(), g['state_%d' % t], rtol=0, atol=STATE_TOL, equal_nan=True)
        env.harvest()
    np.testing.assert_array_equal(env.flags, g['final_flags'])
    np.testing.assert_array_equal(env.lengths, g['final_lengths'])
    tr = env.get_streamlines()
    np.testing.assert_array_equal(tr.lengths, g['sl_lengths'])
    OracleSingleton.clear()


def test_fp16_oracle_tier_inside_step_decides_like_the_fp32_tier():
    """S9 / S13 with the production oracle tier (fp16 tensor cores, what `oracle_precision` defaults to):
    the same episode as the reference fixture, once with the fp32 oracle tier and once with the fp16 one.
    A streamline whose fp32 score never comes within 5e-3 of the 0.5 threshold must end with the same
    flags and length in both; the fp32 run itself is pinned to the reference fixture by the test above."""
    from tests.gpu_helpers import make_gpu_env
    from tests.helpers import oracle_ckpt_for
    from tracktolearn_b200.oracles.oracle import OracleSingleton
    g = load_golden('env_oracle')
    n = len(g['seeds'])
    results = {}
    borderline = np.zeros(n, dtype=bool)
    for prec in ('fp32', 'fp16'):
        OracleSingleton.clear()
        env, _ = make_gpu_env(g, False, True, seeds=g['seeds'], oracle_checkpoint=oracle_ckpt_for(g),
                              oracle_stopping=True, oracle_bonus=10.0, min_length=1.6, oracle_precision=prec)
        assert env._oracle.precision == prec
        env.reset(0, n)
        rewards = np.zeros(n)
        for t in range(int(g['n_steps'])):
            ci = env.continue_idx.copy()
            if len(ci) == 0:
                break
            st, r, done, _ = env.step(g['actions'][t][ci])
            rewards[ci] += r
            if prec == 'fp32' and env.length > env.min_nb_steps:
                sc = env._oracle_scores[:len(ci)].cpu().numpy()
                borderline[ci[np.abs(sc - 0.5) <= 5e-3]] = True
            env.harvest()
        results[prec] = (env.flags.copy(), env.lengths.copy(), rewards)
    OracleSingleton.clear()
    np.testing.assert_array_equal(results['fp32'][0], g['final_flags'])
    clear = ~borderline
    assert clear.mean() > 0.8, clear.mean()
    np.testing.assert_array_equal(results['fp16'][0][clear], results['fp32'][0][clear])
    np.testing.assert_array_equal(results['fp16'][1][clear], results['fp32'][1][clear])
    np.testing.assert_allclose(results['fp16'][2][clear], results['fp32'][2][clear], rtol=0, atol=1e-4)
    assert (results['fp32'][0] & 64).any()        # the ORACLE criterion did fire


def test_empty_and_single_streamline_batches():
    """Edge sizes of the reference protocol: an empty batch (reset(k, k)), a single streamline, and a
    batch that empties while stepping -- no launch with zero rows, shapes as in the reference."""
    from tests.gpu_helpers import make_gpu_env
    g = load_golden('env_noisy')
    env, sub = make_gpu_env(g, True, False, seeds=g['seeds'])
    # empty
    s = env.reset(3, 3)
    assert tuple(s.shape) == (0, env.get_state_size())
    st, r, d, info = env.step(np.zeros((0, 3), dtype=np.float32))
    assert tuple(st.shape) == (0, env.get_state_size()) and len(d) == 0 and len(info['continue_idx']) == 0
    st, ns = env.harvest()
    assert tuple(st.shape) == (0, env.get_state_size()) and len(ns) == 0
    assert len(env.get_streamlines()) == 0
    # one streamline, tracked by hand until it stops; matches the oracle on the same actions
    seeds = np.asarray(env.seeds[5:6])
    ref = O.OracleEnv(sub['sh'], sub['mask'], seeds, meta(g)['vox'], meta(g)['step_mm'], theta=meta(g)['theta'],
                      max_length_mm=meta(g)['max_length'], threshold=meta(g)['threshold'], noisy=True)
    s = env.reset(5, 6)
    s_ref = ref.reset(0, 1)
    np.testing.assert_allclose(s.cpu().numpy(), s_ref, atol=1e-5)
    rs = np.random.RandomState(0)
    a = rs.normal(size=(1, 3)).astype(np.float32)
    steps = 0
    while len(env.continue_idx):
        st, _, d, _ = env.step(a)
        st_r, _, d_r, _ = ref.step(a)
        assert (d == d_r).all()
        np.testing.assert_allclose(st.cpu().numpy(), st_r, atol=1e-5)
        env.harvest()
        ref.harvest()
        steps += 1
        assert steps <= env.max_nb_steps
    assert len(ref.continue_idx) == 0
    t = env.get_streamlines()
    sl, _, fl = ref.get_streamlines()
    assert len(t) == 1 and t.lengths[0] == len(sl[0]) and t.data_per_streamline['flags'][0] == fl[0]
    np.testing.assert_allclose(t.streamlines[0], sl[0], atol=1e-5)
    # stepping an already empty env is a no-op
    st, r, d, info = env.step(np.zeros((0, 3), dtype=np.float32))
    assert tuple(st.shape) == (0, env.get_state_size())
