"""CLI smoke tests mirroring the reference's tests/test_runners.py (--help must work without a GPU)
plus, on the GPU box, an end-to-end ttl_track run on a tiny synthetic subject written to disk."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_help_option():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ttl_track.py'), '--help'],
                         capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ('in_odf', 'in_seed', 'in_mask', 'out_tractogram', '--sh_basis', '--compress', '--save_seeds',
                 '--agent', '--hyperparameters', '--n_actor', '--npv', '--min_length', '--max_length',
                 '--noise', '--fa_map', '--binary_stopping_threshold', '--rng_seed'):
        assert flag in out.stdout, flag


def test_sh_basis_conversion_is_an_involution():
    from tracktolearn_b200.datasets.files import set_sh_order_basis
    rs = np.random.RandomState(0)
    sh = rs.normal(size=(2, 3, 4, 45)).astype(np.float32)
    once = set_sh_order_basis(sh, 'tournier07', target_order=8)
    assert not np.array_equal(once, sh)
    np.testing.assert_array_equal(set_sh_order_basis(once, 'tournier07', target_order=8), sh)
    np.testing.assert_array_equal(once[..., 0], sh[..., 0])        # l = 0 untouched
    # order 6 (28 coefs) -> order 8: zero padded; full basis -> even degrees only
    assert set_sh_order_basis(sh[..., :28], 'descoteaux07', target_order=8).shape[-1] == 45
    full = rs.normal(size=(1, 1, 1, 81)).astype(np.float32)
    assert set_sh_order_basis(full, 'descoteaux07', target_order=8).shape[-1] == 45


@pytest.mark.gpu
@pytest.mark.parametrize('ext', ['trk', 'tck'])
def test_ttl_track_end_to_end(tmp_path, ext):
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.io import nifti
    from tracktolearn_b200.io.streamlines import read_tck, read_trk
    from tracktolearn_b200.runners.ttl_track import main
    shape = (24, 26, 22)
    sub = synthetic.make_subject(shape, seed=5)
    affine = np.diag([1.25, 1.25, 1.25, 1.0])
    affine[:3, 3] = [-10.0, 4.0, 2.5]
    nifti.save(str(tmp_path / 'fodf.nii.gz'), sub['sh'].numpy(), affine)
    nifti.save(str(tmp_path / 'mask.nii.gz'), sub['mask'].numpy(), affine)
    nifti.save(str(tmp_path / 'seed.nii.gz'), synthetic.ellipsoid_mask(shape, frac=0.3).numpy().astype(np.uint8), affine)
    agent = synthetic.write_agent_dir(str(tmp_path / 'agent'), kind='tracking', hidden_dims='128-128-128')
    out = str(tmp_path / ('out.' + ext))
    main([str(tmp_path / 'fodf.nii.gz'), str(tmp_path / 'seed.nii.gz'), str(tmp_path / 'mask.nii.gz'), out,
          '--agent', agent, '--hyperparameters', os.path.join(agent, 'hyperparameters.json'),
          '--n_actor', '500', '--npv', '2', '--min_length', '5', '--max_length', '60', '--save_seeds'])
    data, offsets, hdr = (read_trk if ext == 'trk' else read_tck)(out)
    n = len(offsets) - 1
    assert n > 200
    lens = np.diff(offsets)
    assert lens.min() >= 2
    if ext == 'trk':
        assert hdr['n_count'] == n and np.allclose(hdr['voxel_sizes'], 1.25)
        vox = data / 1.25 - 0.5                      # back to voxel space
    else:
        vox = (data - affine[:3, 3]) / 1.25
    assert vox.min() > -1 and (vox.max(axis=0) < np.asarray(shape)).all()
    seg = np.linalg.norm(np.diff(vox, axis=0), axis=1)
    inner = np.ones(len(vox) - 1, dtype=bool)
    inner[offsets[1:-1] - 1] = False
    # step size rescaled by voxel size: 1.25 / 0.9987237 * 0.75 mm = 0.75096 voxels
    np.testing.assert_allclose(seg[inner], 0.75 / 0.9987237, rtol=2e-4)


@pytest.mark.gpu
def test_ttl_track_compress_flag(tmp_path):
    """`--compress t` (ttl_track.py:223-228 -> tracker.py:103,123-125): same streamlines, fewer points,
    every removed point within t mm = t / voxel_size voxels (compression runs in voxel space, before the
    space change) of the chord that replaced it; end points untouched."""
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.io import nifti
    from tracktolearn_b200.io.streamlines import read_tck
    from tracktolearn_b200.runners.ttl_track import main
    shape = (24, 26, 22)
    sub = synthetic.make_subject(shape, seed=5)
    affine = np.diag([1.25, 1.25, 1.25, 1.0])
    nifti.save(str(tmp_path / 'fodf.nii.gz'), sub['sh'].numpy(), affine)
    nifti.save(str(tmp_path / 'mask.nii.gz'), sub['mask'].numpy(), affine)
    nifti.save(str(tmp_path / 'seed.nii.gz'), synthetic.ellipsoid_mask(shape, frac=0.3).numpy().astype(np.uint8), affine)
    agent = synthetic.write_agent_dir(str(tmp_path / 'agent'), kind='tracking', hidden_dims='128-128-128')
    outs = []
    for extra in ([], ['--compress', '0.1']):
        out = str(tmp_path / ('out%d.tck' % len(extra)))
        main([str(tmp_path / 'fodf.nii.gz'), str(tmp_path / 'seed.nii.gz'), str(tmp_path / 'mask.nii.gz'), out,
              '--agent', agent, '--hyperparameters', os.path.join(agent, 'hyperparameters.json'),
              '--n_actor', '500', '--npv', '1', '--min_length', '5', '--max_length', '60', '--rng_seed', '7'] + extra)
        outs.append(read_tck(out))
    (d0, o0, _), (d1, o1, _) = outs
    assert len(o0) == len(o1) and len(o0) > 100
    assert len(d1) < 0.8 * len(d0)
    for i in range(len(o0) - 1):
        a, b = d0[o0[i]:o0[i + 1]] / 1.25, d1[o1[i]:o1[i + 1]] / 1.25
        np.testing.assert_allclose(a[0], b[0], atol=1e-5)
        np.testing.assert_allclose(a[-1], b[-1], atol=1e-5)
        # every original point lies within the tolerance of the compressed polyline
        seg0, seg1 = b[:-1], b[1:]
        u = seg1 - seg0
        for p in a[:: max(1, len(a) // 8)]:
            t = np.clip(((p - seg0) * u).sum(1) / np.maximum((u * u).sum(1), 1e-12), 0, 1)
            dist = np.linalg.norm(seg0 + t[:, None] * u - p, axis=1).min()
            assert dist <= 0.1 / 1.25 + 1e-3, dist


def test_training_help_option():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'sac_auto_train.py'), '--help'],
                         capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ('path', 'experiment', 'id', '--max_ep', '--log_interval', '--lr', '--gamma', '--alignment_weighting',
                 '--n_actor', '--hidden_dims', '--npv', '--theta', '--min_length', '--max_length', '--step_size',
                 '--noise', '--n_dirs', '--oracle_checkpoint', '--oracle_bonus', '--alpha', '--batch_size',
                 '--replay_size', '--rng_seed'):
        assert flag in out.stdout, flag


@pytest.mark.gpu
def test_sac_auto_train_end_to_end_then_track_with_the_trained_agent(tmp_path):
    """configs[3]'s surface: the sac_auto_train runner trains for two episodes on a tiny synthetic subject,
    writes <path>/model/{hyperparameters.json, last_model_state_actor.pth, last_model_state_critic.pth}
    in the reference's layout (trainers/train.py:151-179, sac_auto_train.py:48-59), and ttl_track.py then
    tracks with that directory."""
    import json
    import torch
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.io import nifti
    from tracktolearn_b200.io.streamlines import read_tck
    from tracktolearn_b200.runners.ttl_track import main as track_main
    from tracktolearn_b200.trainers.sac_auto_train import main as train_main
    shape = (24, 26, 22)
    sub = synthetic.make_subject(shape, seed=5)
    affine = np.diag([1.0, 1.0, 1.0, 1.0])
    nifti.save(str(tmp_path / 'fodf.nii.gz'), sub['sh'].numpy(), affine)
    nifti.save(str(tmp_path / 'mask.nii.gz'), sub['mask'].numpy(), affine)
    nifti.save(str(tmp_path / 'seed.nii.gz'), synthetic.ellipsoid_mask(shape, frac=0.3).numpy().astype(np.uint8), affine)
    exp = train_main([str(tmp_path / 'exp'), 'unit', 'run0', str(tmp_path / 'fodf.nii.gz'), str(tmp_path / 'seed.nii.gz'),
                      str(tmp_path / 'mask.nii.gz'), '--max_ep', '2', '--log_interval', '1', '--n_actor', '96',
                      '--hidden_dims', '64-64-64', '--batch_size', '64', '--replay_size', '8192',
                      '--start_timesteps', '96', '--min_length', '3', '--max_length', '30', '--npv', '1'])
    model = tmp_path / 'exp' / 'model'
    hp = json.load(open(model / 'hyperparameters.json'))
    for key in ('algorithm', 'step_size', 'voxel_size', 'max_angle', 'hidden_dims', 'n_dirs', 'target_sh_order',
                'input_size', 'action_size', 'alpha', 'batch_size', 'replay_size', 'lr', 'gamma', 'n_actor',
                'min_length', 'max_length', 'alignment_weighting', 'binary_stopping_threshold', 'noise'):
        assert key in hp, key
    assert hp['algorithm'] == 'SACAuto' and hp['input_size'] == 615 and isinstance(hp['voxel_size'], str)
    actor = torch.load(model / 'last_model_state_actor.pth', map_location='cpu')
    critic = torch.load(model / 'last_model_state_critic.pth', map_location='cpu')
    assert tuple(actor['layers.0.weight'].shape) == (64, 615) and tuple(critic['q1.0.weight'].shape) == (64, 618)
    episodes = [e for e in exp.training_log if 'avg_length' in e]
    assert len(episodes) == 2 and all(e['avg_length'] > 1 for e in episodes)
    assert any(e['losses'] for e in episodes)            # updates did run
    out = str(tmp_path / 'trained.tck')
    track_main([str(tmp_path / 'fodf.nii.gz'), str(tmp_path / 'seed.nii.gz'), str(tmp_path / 'mask.nii.gz'), out,
                '--agent', str(model), '--hyperparameters', str(model / 'hyperparameters.json'),
                '--n_actor', '200', '--npv', '1', '--min_length', '0', '--max_length', '60'])
    data, offsets, _ = read_tck(out)
    assert len(offsets) - 1 > 50
