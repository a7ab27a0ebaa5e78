#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's tests pin no numbers and its third-party stack is not
installable here (F9), so this script imports ``/root/reference/TrackToLearn`` with
the stand-ins of ``_reference_stubs.py`` for the absent packages and records what the
reference's unmodified ``TrackingEnvironment`` / ``NoisyTrackingEnvironment``
(reset / step / harvest / get_streamlines, _format_state, _compute_stopping_flags,
PeaksAlignmentReward), ``MaxEntropyActor`` and ``TransformerOracle`` compute on seeded
synthetic inputs.  The fixtures are what ``oracle/`` is pinned against
(tests/test_oracle_golden.py) and what the CUDA path is checked against on the GPU box
(tests/test_*_gpu.py) -- /root/reference does not exist there.

numpy caveat: the reference pins numpy 1.23 (install.sh:29); this container has numpy 2.x
whose NEP-50 promotion turns ``float32_array * np.float64_scalar`` into float64.  For the
plain ``TrackingEnvironment`` cases we therefore hand ``convert_length_mm2vox``'s result
over as a Python float (a weak scalar), which reproduces numpy 1.23's float32 arithmetic
in ``_format_actions``.  The Noisy environment (every ``ttl_track`` run) is float64 under
both numpy versions and is run unpatched.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _reference_stubs  # noqa: E402

_reference_stubs.install()

from TrackToLearn.datasets.utils import MRIDataVolume  # noqa: E402
from TrackToLearn.environments import env as ref_env_mod  # noqa: E402
from TrackToLearn.environments.noisy_tracking_env import NoisyTrackingEnvironment  # noqa: E402
from TrackToLearn.environments.tracking_env import TrackingEnvironment  # noqa: E402
from TrackToLearn.environments.stopping_criteria import StoppingFlags  # noqa: E402

from tracktolearn_b200 import synthetic  # noqa: E402

assert int(np.__version__.split('.')[0]) >= 2


def make_env(cls, shape, vox, npv, step_mm, theta=30., n_dirs=100, max_length=200.,
             min_length=20., compute_reward=False, threshold=0.1, np_seed=1337,
             vol_seed=1234, weak_step=False, noise=0.0, seed_frac=0.26, oracle_checkpoint=None,
             oracle_stopping=False, oracle_bonus=0.0):
    sub = synthetic.make_subject(shape, seed=vol_seed)
    affine = np.diag([vox, vox, vox, 1.0])
    vol = MRIDataVolume(sub['sh'].numpy(), affine)
    mask = MRIDataVolume(sub['mask'].numpy().astype(np.float64), affine)
    # seed inside the mask (not on its shell) so that episodes last more than a few steps
    seed_mask = synthetic.ellipsoid_mask(shape, frac=seed_frac).numpy()
    seed = MRIDataVolume(seed_mask.astype(np.float64), affine)
    peaks = MRIDataVolume(sub['peaks'].numpy(), affine)
    dto = {
        'n_dirs': n_dirs, 'theta': theta, 'npv': npv, 'binary_stopping_threshold': threshold,
        'step_size': step_mm, 'min_length': min_length, 'max_length': max_length,
        'oracle_checkpoint': oracle_checkpoint, 'oracle_stopping_criterion': oracle_stopping,
        'scoring_data': None,
        'compute_reward': compute_reward, 'alignment_weighting': 1.0, 'oracle_bonus': oracle_bonus,
        'rng': np.random.RandomState(np_seed), 'device': torch.device('cpu'),
        'target_sh_order': 8, 'noise': noise, 'fa_map': None,
    }
    orig = ref_env_mod.convert_length_mm2vox
    if weak_step:
        ref_env_mod.convert_length_mm2vox = lambda mm, aff: float(orig(mm, aff))
    try:
        np.random.seed(np_seed)
        env = cls((vol, mask, seed, peaks, affine), 'testing', dto)
    finally:
        ref_env_mod.convert_length_mm2vox = orig
    return env, sub


def make_actions(n, n_steps, rng, kink_every=7):
    """[T, N, 3] float32 actions per GLOBAL row: smooth random walk on the sphere, with a few
    sharp kinks (curvature stops) and one zero action (NaN direction)."""
    a = rng.normal(size=(n, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    out = np.zeros((n_steps, n, 3), dtype=np.float32)
    for t in range(n_steps):
        a = a + 0.12 * rng.normal(size=(n, 3))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        out[t] = (a * rng.uniform(0.2, 1.0, size=(n, 1))).astype(np.float32)
        if t > 2 and t % kink_every == 0:
            k = rng.randint(0, n, size=max(1, n // 16))
            out[t, k] = rng.normal(size=(len(k), 3)).astype(np.float32)
    return out


def run_episode(env, n, actions, keep_states=(0, 1, 2, 3, 4, 9, 19)):
    rec = {}
    state = env.reset(0, n)
    rec['seeds'] = np.asarray(env.seeds[:n], dtype=np.float64)
    rec['state_reset'] = state.cpu().numpy()
    T = actions.shape[0]
    alive_counts, cont_idx, dones_l, rewards, new_pts, flags_l, states = [], [], [], [], [], [], {}
    t = 0
    done = np.array([False])
    while not np.all(done) and t < T:
        ci = env.continue_idx.copy()
        if t == 2:
            actions[t, ci[0]] = 0.0  # zero action -> NaN direction -> NaN position
        st, reward, done, info = env.step(actions[t][ci])
        alive_counts.append(len(ci))
        cont_idx.append(ci.astype(np.int32))
        dones_l.append(np.asarray(done, dtype=np.uint8))
        rewards.append(np.asarray(reward, dtype=np.float64)[:len(ci)] if len(reward) != len(ci)
                       else np.asarray(reward, dtype=np.float64))
        new_pts.append(env.streamlines[ci, env.length - 1].copy())
        flags_l.append(env.flags[ci].astype(np.int32))
        if t in keep_states:
            states[t] = st.cpu().numpy()
        hst, _ = env.harvest()
        if t in keep_states:
            rec['harvest_state_%d' % t] = hst.cpu().numpy()
        t += 1
    assert np.all(done), 'episode did not finish in %d steps' % T
    rec['n_steps'] = np.int32(t)
    rec['alive_counts'] = np.asarray(alive_counts, dtype=np.int32)
    rec['continue_idx'] = np.concatenate(cont_idx)
    rec['dones'] = np.concatenate(dones_l)
    rec['rewards'] = np.concatenate(rewards)
    rec['new_points'] = np.concatenate(new_pts).astype(np.float32)
    rec['step_flags'] = np.concatenate(flags_l)
    for k, v in states.items():
        rec['state_%d' % k] = v
    rec['final_flags'] = env.flags.astype(np.int32)
    rec['final_lengths'] = env.lengths.astype(np.int32)
    tr = env.get_streamlines()
    rec['sl_lengths'] = np.asarray([len(s) for s in tr.streamlines], dtype=np.int32)
    rec['sl_points'] = np.concatenate(tr.streamlines).astype(np.float32)
    rec['actions'] = actions[:t]
    return rec


def case_env(name, cls, weak_step, compute_reward, shape=(20, 22, 18), vox=1.0, step_mm=0.75,
             n=48, max_length=13.6, theta=30.0):
    env, sub = make_env(cls, shape, vox, npv=1, step_mm=step_mm, max_length=max_length,
                        compute_reward=compute_reward, weak_step=weak_step, theta=theta)
    np.random.RandomState(7).shuffle(env.seeds)
    n = min(n, len(env.seeds))
    rng = np.random.RandomState(99)
    actions = make_actions(n, env.max_nb_steps + 2, rng)
    rec = run_episode(env, n, actions)
    rec['meta_shape'] = np.asarray(shape, dtype=np.int32)
    rec['meta'] = np.asarray([vox, step_mm, theta, max_length, 0.1, float(env.max_nb_steps),
                              float(env.step_size)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    fl = rec['final_flags']
    print(name, 'n=%d steps=%d' % (n, rec['n_steps']),
          'flags: mask=%d length=%d curv=%d' % ((fl & 1 > 0).sum(), (fl & 2 > 0).sum(), (fl & 4 > 0).sum()),
          'nan_points=%d' % np.isnan(rec['new_points']).any(axis=1).sum())


def case_env_oracle(name='env_oracle'):
    """TrackingEnvironment with the oracle stopping criterion and the sparse oracle bonus
    (stopping_criteria.py:85-154, oracle_reward.py:10-93) driven by a 1-layer TransformerOracle
    whose head is rescaled so that scores straddle 0.5."""
    import tempfile
    from TrackToLearn.oracles.oracle import OracleSingleton
    from TrackToLearn.oracles.transformer_oracle import TransformerOracle
    ck = synthetic.oracle_checkpoint(n_head=4, n_layers=1, input_size=384, seed=31)
    model = TransformerOracle.load_from_checkpoint(
        {'hyper_parameters': ck['hyper_parameters'],
         'state_dict': {k: v.clone() for k, v in ck['state_dict'].items()}})
    rng = np.random.RandomState(8)
    probe = synthetic.random_streamlines(64, rng, min_pts=4, max_pts=20)
    feats = np.diff(np.stack(_reference_stubs.set_number_of_points(probe, 128)), axis=1).astype(np.float32)
    with torch.no_grad():
        hidden = model.bert(model.pos_encoding(model.embedding(torch.cat(
            (model.cls_token.repeat(64, 1, 1), torch.from_numpy(feats)), dim=1)) * np.sqrt(32.0)))[:, 0]
        logit = (hidden @ ck['state_dict']['head.weight'].t())[:, 0]
    scale = 6.0 / float(logit.std())
    head_w = ck['state_dict']['head.weight'] * scale
    head_b = torch.tensor([-float((logit * scale).median())])
    ck['state_dict']['head.weight'] = head_w
    ck['state_dict']['head.bias'] = head_b
    path = os.path.join(tempfile.mkdtemp(), 'oracle.ckpt')
    torch.save(ck, path)
    OracleSingleton._self = None
    shape, vox, step_mm = (20, 22, 18), 1.0, 0.75
    env, sub = make_env(TrackingEnvironment, shape, vox, npv=1, step_mm=step_mm, max_length=13.6,
                        min_length=1.6, compute_reward=True, weak_step=True, oracle_checkpoint=path,
                        oracle_stopping=True, oracle_bonus=10.0)
    np.random.RandomState(7).shuffle(env.seeds)
    n = min(48, len(env.seeds))
    actions = make_actions(n, env.max_nb_steps + 2, np.random.RandomState(99))
    rec = run_episode(env, n, actions, keep_states=(0, 12))
    rec['meta_shape'] = np.asarray(shape, dtype=np.int32)
    rec['meta'] = np.asarray([vox, step_mm, 30.0, 13.6, 0.1, float(env.max_nb_steps), float(env.step_size)],
                             dtype=np.float64)
    rec['head_w'] = head_w.numpy()
    rec['head_b'] = head_b.numpy()
    rec['min_nb_steps'] = np.int32(env.min_nb_steps)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    fl = rec['final_flags']
    print(name, 'n=%d steps=%d' % (n, rec['n_steps']),
          'flags: mask=%d length=%d curv=%d oracle=%d' % ((fl & 1 > 0).sum(), (fl & 2 > 0).sum(),
                                                          (fl & 4 > 0).sum(), (fl & 64 > 0).sum()),
          'bonus rows=%d' % int((rec['rewards'] > 5).sum()))
    OracleSingleton._self = None


def case_sac_update(name='sac_update'):
    """Two SACAuto.update steps of the reference (sac_auto.py:139-250) on a fixed batch, small nets."""
    from TrackToLearn.algorithms.sac_auto import SACAuto
    torch.manual_seed(123)
    alg = SACAuto(24, 3, '16-12', lr=3e-4, gamma=0.95, alpha=0.2, n_actors=8, batch_size=16, replay_size=8,
                  rng=np.random.RandomState(0), device=torch.device('cpu'))
    rec = {}
    for k, v in alg.agent.actor.state_dict().items():
        rec['actor0.' + k] = v.numpy().copy()
    for k, v in alg.agent.critic.state_dict().items():
        rec['critic0.' + k] = v.numpy().copy()
    g = torch.Generator().manual_seed(5)
    batch = (torch.randn((16, 24), generator=g), torch.rand((16, 3), generator=g) * 2 - 1,
             torch.randn((16, 24), generator=g), torch.rand((16,), generator=g),
             (torch.rand((16,), generator=g) > 0.3).float())
    for i, name_ in enumerate(('state', 'action', 'next_state', 'reward', 'not_done')):
        rec['batch.' + name_] = batch[i].numpy()
    torch.manual_seed(77)
    for _ in range(2):
        alg.update(batch)
    for k, v in alg.agent.actor.state_dict().items():
        rec['actor2.' + k] = v.numpy().copy()
    for k, v in alg.agent.critic.state_dict().items():
        rec['critic2.' + k] = v.numpy().copy()
    for k, v in alg.target.critic.state_dict().items():
        rec['target_critic2.' + k] = v.numpy().copy()
    rec['log_alpha2'] = alg.log_alpha.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    print(name, 'log_alpha', rec['log_alpha2'])


def case_edges(name='edges'):
    """_format_state / stopping flags / reward on crafted streamlines: outside the volume,
    negative, on the lattice, NaN/inf, zero-length segments."""
    shape = (12, 10, 9)
    env, sub = make_env(TrackingEnvironment, shape, 1.25, npv=1, step_mm=0.9, max_length=20.,
                        compute_reward=True, weak_step=True)
    rng = np.random.RandomState(5)
    n, L = 64, 6
    pts = rng.uniform(1.0, 9.0, size=(n, 1, 3)) + np.cumsum(
        rng.normal(scale=0.45, size=(n, L, 3)), axis=1)
    pts = pts.astype(np.float32)
    pts[0, -1] = [3.0, 4.0, 5.0]            # exact lattice point
    pts[1, -1] = [0.0, 0.0, 0.0]
    pts[2, -1] = [11.0, 9.0, 8.0]           # last voxel
    pts[3, -1] = [11.4, 9.2, 8.49]          # beyond last voxel centre
    pts[4, -1] = [-0.5, 4.0, 4.0]
    pts[5, -1] = [np.nan, 4.0, 4.0]
    pts[6, -1] = [np.inf, 4.0, 4.0]
    pts[7, -1] = [-np.inf, 2.0, 3.0]
    pts[8, -1] = pts[8, -2]                 # zero-length last segment
    pts[9, -2] = pts[9, -3]                 # zero-length previous segment
    pts[10, -1] = [5.5, 5.5, 4.5]           # mask coords exactly on lattice after -0.5
    pts[11, -1] = [0.49, 0.5, 0.51]         # mask coords straddle 0
    pts[12, :, :] = np.linspace(2, 6, L)[:, None]   # perfectly straight
    pts[13, -1] = 2 * pts[13, -2] - pts[13, -3]     # collinear continuation (dot ~ 1)
    pts[14, -1] = pts[14, -3]                        # 180-degree reversal (dot ~ -1)
    rec = {'points': pts}
    with np.errstate(all='ignore'):
        for Lk in (1, 2, 3, L):
            sub_pts = pts[:, -Lk:] if Lk > 1 else pts[:, -1:]
            rec['state_L%d' % Lk] = env._format_state(sub_pts).cpu().numpy()
            stop, flags = env._is_stopping(sub_pts)
            rec['stop_L%d' % Lk] = stop.astype(np.uint8)
            rec['flags_L%d' % Lk] = flags.astype(np.int32)
            r, _ = env.reward_function(sub_pts, np.zeros(n, dtype=bool))
            rec['reward_L%d' % Lk] = np.asarray(r, dtype=np.float64)
        mask_crit = env.stopping_criteria[StoppingFlags.STOPPING_MASK]
        from scipy.ndimage import map_coordinates
        rec['mask_values'] = map_coordinates(mask_crit.mask, pts[:, -1, :].T - 0.5, prefilter=False)
    rec['meta_shape'] = np.asarray(shape, dtype=np.int32)
    rec['meta'] = np.asarray([1.25, 0.9, 30.0, 20.0, 0.1, float(env.max_nb_steps),
                              float(env.step_size)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    print(name, 'stops L6:', int(rec['stop_L%d' % L].sum()), 'flags:', np.unique(rec['flags_L%d' % L]))


def case_actor(name='actor'):
    """Reference MaxEntropyActor (offpolicy.py:61-140) forward, deterministic and sampled."""
    from TrackToLearn.algorithms.shared.offpolicy import MaxEntropyActor
    hidden = '96-64-80'
    sd = synthetic.actor_state_dict(615, hidden, seed=4321)
    actor = MaxEntropyActor(615, 3, hidden)
    actor.load_state_dict(sd)
    actor.eval()
    g = torch.Generator().manual_seed(7)
    state = torch.randn((37, 615), generator=g)
    state[:, 315:] *= 0.5
    with torch.no_grad():
        p = actor.layers(state)
        torch.manual_seed(11)
        a0, lp0 = actor(state, 0.0)
        torch.manual_seed(11)
        eps = torch.randn((37, 3))       # Normal.rsample draws eps = randn(shape) first
        torch.manual_seed(11)
        a1, lp1 = actor(state, 1.0)
    rec = {'state': state.numpy(), 'pre': p.numpy(), 'action_det': a0.numpy(),
           'logp_det': lp0.numpy(), 'eps': eps.numpy(), 'action_prob1': a1.numpy(),
           'logp_prob1': lp1.numpy(), 'hidden': np.asarray([96, 64, 80], dtype=np.int32),
           'seed': np.int32(4321)}
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    print(name, 'max|a|', float(a0.abs().max()))


def case_oracle_net(name='oracle_net'):
    """Reference TransformerOracle.forward (transformer_oracle.py:77-92) in fp32 on CPU and
    OracleSingleton.predict's resample+diff front end (oracle.py:52-54)."""
    from TrackToLearn.oracles.transformer_oracle import TransformerOracle
    ck = synthetic.oracle_checkpoint(n_head=4, n_layers=2, input_size=384, seed=2222)
    model = TransformerOracle.load_from_checkpoint(
        {'hyper_parameters': ck['hyper_parameters'],
         'state_dict': {k: v.clone() for k, v in ck['state_dict'].items()}})
    model.eval()
    rng = np.random.RandomState(3)
    sl = synthetic.random_streamlines(24, rng, min_pts=2, max_pts=90)
    sl[0] = sl[0][:2]
    sl[1] = sl[1][:3]
    res = _reference_stubs.set_number_of_points(sl, 128)
    dirs = np.diff(np.stack(res), axis=1).astype(np.float32)
    with torch.no_grad():
        scores = model(torch.from_numpy(dirs)).numpy()
    rec = {'sl_lengths': np.asarray([len(s) for s in sl], dtype=np.int32),
           'sl_points': np.concatenate(sl).astype(np.float32),
           'dirs': dirs, 'scores': scores.astype(np.float32),
           'hp': np.asarray([4, 2, 384, 2222], dtype=np.int32)}
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    print(name, 'scores', scores[:5])


if __name__ == '__main__':
    with np.errstate(all='ignore'):
        case_env('env_noisy', NoisyTrackingEnvironment, weak_step=False, compute_reward=False)
        case_env('env_plain_reward', TrackingEnvironment, weak_step=True, compute_reward=True,
                 shape=(18, 20, 22), vox=1.25, step_mm=0.9375, theta=40.0, max_length=17.0)
        case_edges()
        case_env_oracle()
    case_actor()
    case_oracle_net()
    case_sac_update()
