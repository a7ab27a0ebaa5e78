"""The CPU oracle (oracle/ttl_oracle.py) against fixtures recorded from the reference's own
code (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import ttl_oracle as O
from tests.helpers import load_golden, meta, split_by_counts, subject_for
from tracktolearn_b200 import synthetic

STATE_TOL = 1e-5


def _make_env(g, noisy, compute_reward):
    sub = subject_for(g)
    m = meta(g)
    env = O.OracleEnv(sub['sh'], sub['mask'], g['seeds'], m['vox'], m['step_mm'], theta=m['theta'],
                      max_length_mm=m['max_length'], threshold=m['threshold'], peaks=sub['peaks'],
                      compute_reward=compute_reward, noisy=noisy, noise=0.0)
    assert env.max_nb_steps == m['max_nb_steps']
    assert env.step_size == pytest.approx(m['step_vox'], abs=0)
    return env


@pytest.mark.parametrize('name,noisy,reward', [('env_noisy', True, False),
                                               ('env_plain_reward', False, True)])
def test_env_episode_matches_reference(name, noisy, reward):
    g = load_golden(name)
    env = _make_env(g, noisy, reward)
    n = len(g['seeds'])
    state = env.reset(0, n)
    np.testing.assert_allclose(state, g['state_reset'], rtol=0, atol=STATE_TOL)
    counts = g['alive_counts']
    ci_g = split_by_counts(g['continue_idx'], counts)
    dones_g = split_by_counts(g['dones'], counts)
    pts_g = split_by_counts(g['new_points'], counts)
    flags_g = split_by_counts(g['step_flags'], counts)
    rew_g = split_by_counts(g['rewards'], counts)
    for t in range(int(g['n_steps'])):
        ci = env.continue_idx.copy()
        np.testing.assert_array_equal(ci, ci_g[t])
        st, r, done, _ = env.step(g['actions'][t][ci])
        np.testing.assert_array_equal(done.astype(np.uint8), dones_g[t])
        np.testing.assert_allclose(env.streamlines[ci, env.length - 1], pts_g[t], rtol=0, atol=1e-6,
                                   equal_nan=True)
        np.testing.assert_array_equal(env.flags[ci], flags_g[t])
        if reward:
            np.testing.assert_allclose(r, rew_g[t], rtol=0, atol=1e-6)
        if 'state_%d' % t in g:
            np.testing.assert_allclose(st, g['state_%d' % t], rtol=0, atol=STATE_TOL, equal_nan=True)
        hs, _ = env.harvest()
        if 'harvest_state_%d' % t in g:
            np.testing.assert_allclose(hs, g['harvest_state_%d' % t], rtol=0, atol=STATE_TOL,
                                       equal_nan=True)
    assert len(env.continue_idx) == 0
    np.testing.assert_array_equal(env.flags, g['final_flags'])
    np.testing.assert_array_equal(env.lengths, g['final_lengths'])
    sl, seeds, flags = env.get_streamlines()
    np.testing.assert_array_equal([len(s) for s in sl], g['sl_lengths'])
    np.testing.assert_allclose(np.concatenate(sl), g['sl_points'], rtol=0, atol=1e-6, equal_nan=True)


def test_edges_state_flags_reward():
    g = load_golden('edges')
    sub = subject_for(g)
    m = meta(g)
    env = O.OracleEnv(sub['sh'], sub['mask'], np.zeros((1, 3)), m['vox'], m['step_mm'],
                      theta=m['theta'], max_length_mm=m['max_length'], peaks=sub['peaks'],
                      compute_reward=True)
    pts = g['points']
    L = pts.shape[1]
    for Lk in (1, 2, 3, L):
        sp = pts[:, -Lk:]
        np.testing.assert_allclose(env._format_state(sp), g['state_L%d' % Lk], rtol=0, atol=STATE_TOL,
                                   equal_nan=True)
        stop, flags = env._is_stopping(sp)
        np.testing.assert_array_equal(stop.astype(np.uint8), g['stop_L%d' % Lk])
        np.testing.assert_array_equal(flags, g['flags_L%d' % Lk])
        r = O.peaks_alignment_reward(sub['peaks'], sp).astype(np.float64)
        np.testing.assert_allclose(r, g['reward_L%d' % Lk], rtol=0, atol=1e-6)
    # the tap-by-tap restatement the CUDA kernel follows == real scipy
    coef = env.mask_criterion.mask
    vals = np.array([O.spline_mask_value_restated(coef, p) for p in pts[:, -1]])
    np.testing.assert_allclose(vals, g['mask_values'], rtol=0, atol=1e-13)
    rng = np.random.RandomState(0)
    rnd = rng.uniform(-1.5, np.array(coef.shape) + 0.5, size=(400, 3)).astype(np.float32)
    ref = env.mask_criterion.values(rnd[:, None, :])
    mine = np.array([O.spline_mask_value_restated(coef, p) for p in rnd])
    np.testing.assert_allclose(mine, ref, rtol=0, atol=1e-13)


def test_actor_matches_reference():
    g = load_golden('actor')
    hidden = '-'.join(str(int(h)) for h in g['hidden'])
    sd = {k: v.numpy() for k, v in synthetic.actor_state_dict(615, hidden, seed=int(g['seed'])).items()}
    a, lp, pre = O.actor_forward(sd, g['state'], 0.0)
    np.testing.assert_allclose(pre, g['pre'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a, g['action_det'], rtol=1e-5, atol=1e-6)
    a1, lp1, _ = O.actor_forward(sd, g['state'], 1.0, eps=g['eps'])
    np.testing.assert_allclose(a1, g['action_prob1'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp1, g['logp_prob1'], rtol=1e-4, atol=1e-4)


def test_oracle_net_matches_reference():
    g = load_golden('oracle_net')
    n_head, n_layers, input_size, seed = [int(v) for v in g['hp']]
    ck = synthetic.oracle_checkpoint(n_head=n_head, n_layers=n_layers, input_size=input_size, seed=seed)
    sl = split_by_counts(g['sl_points'], g['sl_lengths'])
    dirs = O.oracle_features(sl)
    np.testing.assert_allclose(dirs, g['dirs'], rtol=0, atol=1e-6)
    scores = O.transformer_oracle_forward(ck, dirs)
    np.testing.assert_allclose(scores, g['scores'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(O.oracle_predict(ck, sl, batch_size=7), g['scores'], rtol=0, atol=2e-5)


def test_seeds_restatement_matches_loop():
    mask = synthetic.ellipsoid_mask((7, 6, 5), frac=0.4).numpy()
    rs = np.random.RandomState(3)
    fast = synthetic.seeds_from_mask(mask, 3, rs)
    np.random.seed(3)
    where = np.argwhere(mask)
    slow = np.asarray([s + np.random.random(3) - .5 for _ in range(3) for s in where])
    np.testing.assert_array_equal(fast, slow)


def test_env_with_oracle_criterion_and_bonus_matches_reference():
    """OracleStoppingCriterion + OracleReward (sparse bonus 10) inside step()."""
    from tests.helpers import oracle_ckpt_for
    g = load_golden('env_oracle')
    sub = subject_for(g)
    m = meta(g)
    ck = oracle_ckpt_for(g)
    env = O.OracleEnv(sub['sh'], sub['mask'], g['seeds'], m['vox'], m['step_mm'], theta=m['theta'],
                      max_length_mm=m['max_length'], min_length_mm=1.6, peaks=sub['peaks'],
                      compute_reward=True, noisy=False, oracle_ckpt=ck, oracle_stopping=True, oracle_bonus=10.0)
    assert env.min_nb_steps == int(g['min_nb_steps'])
    n = len(g['seeds'])
    env.reset(0, n)
    counts = g['alive_counts']
    dones_g = split_by_counts(g['dones'], counts)
    flags_g = split_by_counts(g['step_flags'], counts)
    rew_g = split_by_counts(g['rewards'], counts)
    for t in range(int(g['n_steps'])):
        ci = env.continue_idx.copy()
        st, r, done, _ = env.step(g['actions'][t][ci])
        np.testing.assert_array_equal(done.astype(np.uint8), dones_g[t])
        np.testing.assert_array_equal(env.flags[ci], flags_g[t])
        np.testing.assert_allclose(r, rew_g[t], rtol=0, atol=1e-5)
        env.harvest()
    np.testing.assert_array_equal(env.flags, g['final_flags'])
    assert (np.asarray(g['final_flags']) & O.ORACLE).any()
