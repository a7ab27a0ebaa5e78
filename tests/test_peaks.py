"""Load-time peak extraction (environments/env.py:405-432): the oracle's
restatement of scilpy get_maximas / dipy peak_directions on CPU, the kernel against it on the GPU."""
import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O
from tracktolearn_b200 import synthetic
from tracktolearn_b200.datasets.peaks import sh_to_sf_matrix
from tracktolearn_b200.datasets.sphere import hemisphere


def test_sphere_and_basis():
    v, e, nb = hemisphere(3)
    assert v.shape == (321, 3) and np.allclose(np.linalg.norm(v, axis=1), 1.0)
    # one representative per antipodal pair: no two directions are opposite or equal
    g = np.abs(v @ v.T)
    np.fill_diagonal(g, 0.0)
    assert g.max() < np.cos(np.deg2rad(7.0))
    # every vertex has 5 or 6 neighbours, the table is symmetric
    deg = (nb >= 0).sum(1)
    assert set(deg.tolist()) <= {5, 6}
    for a in range(len(nb)):
        for b in nb[a][nb[a] >= 0]:
            assert a in nb[b]
    assert len(e) == deg.sum() // 2
    # the product's basis matrix equals the oracle's scipy-based one
    B = sh_to_sf_matrix(v, 8)
    np.testing.assert_allclose(B, O.sh_basis_matrix(v, 8), atol=1e-13)
    # SH -> SF -> least-squares SH is the identity on band-limited functions
    rs = np.random.RandomState(0)
    c = rs.normal(size=45)
    back = np.linalg.lstsq(B, B @ c, rcond=None)[0]
    np.testing.assert_allclose(back, c, atol=1e-10)


def test_oracle_peaks_on_known_fields():
    v, e, _ = hemisphere(3)
    B = O.sh_basis_matrix(v, 8)
    # single fibre along a sphere vertex: the peak is that vertex, value-normalised to length 1
    for k in (0, 17, 200):
        u = v[k]
        sh = (synthetic.real_sh_basis(torch.from_numpy(u[None]), 8)[0] * synthetic._zonal_response(8)).numpy()
        pk = O.peaks_from_sh(sh[None].astype(np.float32), v, e, B)[0].reshape(5, 3)
        assert abs(abs(float(pk[0] @ u)) - 1.0) < 1e-6
        assert np.allclose(pk[1:], 0.0)
    # 90-degree crossing with unequal weights: two peaks, second scaled by its relative height
    u1, u2 = np.asarray([1.0, 0, 0]), np.asarray([0, 1.0, 0])
    Y = synthetic.real_sh_basis(torch.from_numpy(np.stack([u1, u2])), 8).numpy() * synthetic._zonal_response(8).numpy()
    sh = (1.0 * Y[0] + 0.6 * Y[1]).astype(np.float32)
    pk = O.peaks_from_sh(sh[None], v, e, B)[0].reshape(5, 3)
    n = np.linalg.norm(pk, axis=1)
    assert abs(n[0] - 1.0) < 1e-6 and 0.4 < n[1] < 0.8 and np.allclose(n[2:], 0.0)
    assert abs(pk[0] @ u1) / n[0] > np.cos(np.deg2rad(6)) and abs(pk[1] @ u2) / n[1] > np.cos(np.deg2rad(6))
    # all-zero voxel and a voxel whose coefficients cancel in the sum: no peaks
    assert not O.peaks_from_sh(np.zeros((1, 45), np.float32), v, e, B).any()


@pytest.mark.gpu
def test_peaks_kernel_matches_oracle():
    from tracktolearn_b200.datasets.peaks import compute_peaks
    shape = (14, 12, 10)
    sub = synthetic.make_subject(shape, seed=21)
    sh = sub['sh'].numpy()
    v, e, _ = hemisphere(3)
    want = O.peaks_from_sh(sh, v, e, O.sh_basis_matrix(v, 8))
    got = compute_peaks(sh, device='cuda:0').cpu().numpy()
    assert got.shape == want.shape == shape + (15,)
    # identical peak sets except where two candidates tie to the last bits of a double (the matrix
    # product is summed in a different order): allow a handful of voxels
    same = np.isclose(got, want, rtol=0, atol=1e-6).all(-1)
    assert same.mean() > 0.995, same.mean()
    assert (np.abs(got).sum(-1) > 0).sum() == (np.abs(want).sum(-1) > 0).sum()
    # the first peak follows the analytic fibre field of the synthetic subject
    ana = sub['peaks'].numpy()[..., :3]
    m = np.linalg.norm(ana, axis=-1) > 0
    a, b = got[..., :3][m], ana[m]
    cos = np.abs((a * b).sum(1)) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    assert np.degrees(np.arccos(np.clip(cos, 0, 1))).mean() < 6.0


@pytest.mark.gpu
def test_from_files_computes_peaks_for_the_alignment_reward(tmp_path):
    """env.from_files with compute_reward (env.py:311-347,405-432): the peaks volume is extracted at load
    time and the alignment reward of a step equals the oracle's on the same peaks."""
    from tracktolearn_b200.environments import TrackingEnvironment
    from tracktolearn_b200.io import nifti
    shape = (16, 14, 12)
    sub = synthetic.make_subject(shape, seed=8)
    affine = np.diag([1.0, 1.0, 1.0, 1.0])
    nifti.save(str(tmp_path / 'fodf.nii.gz'), sub['sh'].numpy(), affine)
    nifti.save(str(tmp_path / 'mask.nii.gz'), sub['mask'].numpy(), affine)
    nifti.save(str(tmp_path / 'seed.nii.gz'), sub['mask'].numpy(), affine)
    dto = {'n_dirs': 100, 'theta': 30.0, 'npv': 1, 'binary_stopping_threshold': 0.1, 'step_size': 0.75,
           'min_length': 1.0, 'max_length': 30.0, 'oracle_checkpoint': None, 'oracle_stopping_criterion': False,
           'scoring_data': None, 'compute_reward': True, 'alignment_weighting': 1.0, 'oracle_bonus': 0.0,
           'rng': np.random.RandomState(1), 'device': torch.device('cuda:0'), 'target_sh_order': 8,
           'in_odf': str(tmp_path / 'fodf.nii.gz'), 'in_seed': str(tmp_path / 'seed.nii.gz'),
           'in_mask': str(tmp_path / 'mask.nii.gz'), 'sh_basis': 'descoteaux07', 'reference': str(tmp_path / 'fodf.nii.gz')}
    env = TrackingEnvironment.from_files(dto)
    assert env.peaks is not None and tuple(env.peaks.data.shape) == shape + (15,)
    v, e, _ = hemisphere(3)
    want = O.peaks_from_sh(sub['sh'].numpy(), v, e, O.sh_basis_matrix(v, 8))
    pk = np.asarray(env.peaks.data)
    assert np.isclose(pk, want, atol=1e-6).all(-1).mean() > 0.995
    n = min(200, len(env.seeds))
    env.reset(0, n)
    rs = np.random.RandomState(2)
    a = rs.normal(size=(n, 3)).astype(np.float32)
    _, r1, _, _ = env.step(a)
    env.harvest()
    ci = env.continue_idx
    _, r2, _, _ = env.step(rs.normal(size=(len(ci), 3)).astype(np.float32))
    pts = env._batch.points[ci, :3].cpu().numpy()
    ref = O.peaks_alignment_reward(pk, pts)
    np.testing.assert_allclose(r2, ref, atol=2e-6)
    assert np.abs(r2).max() > 0.1


def test_reference_sphere_is_used_when_dipy_is_importable(monkeypatch):
    """environments/env.py:412-415 evaluates peaks on dipy's repulsion724 hemisphere.  dipy is not installed
    here, so stand a minimal `dipy` in and check that the loader takes ITS vertices and edges."""
    import sys
    import types
    from tracktolearn_b200.datasets import sphere
    assert sphere.reference_hemisphere() is None or len(sphere.reference_hemisphere()[0]) == 362
    v, e, nb = sphere.hemisphere(1)

    class FakeHemi(object):
        def __init__(self, vertices, edges):
            self.vertices, self.edges = vertices, edges

        @classmethod
        def from_sphere(cls, s):
            return cls(s.vertices, s.edges)

        def subdivide(self, n):
            assert n == 0
            return self
    dipy = types.ModuleType('dipy')
    data = types.ModuleType('dipy.data')
    core = types.ModuleType('dipy.core')
    core_sphere = types.ModuleType('dipy.core.sphere')
    asked = []

    def get_sphere(name=None):
        asked.append(name)
        return FakeHemi(v, e)
    data.get_sphere = get_sphere
    core_sphere.HemiSphere = FakeHemi
    for name, mod in (('dipy', dipy), ('dipy.data', data), ('dipy.core', core), ('dipy.core.sphere', core_sphere)):
        monkeypatch.setitem(sys.modules, name, mod)
    got = sphere.reference_hemisphere()
    assert asked == ['repulsion724']
    np.testing.assert_array_equal(got[0], v)
    np.testing.assert_array_equal(got[1], e)
    np.testing.assert_array_equal(got[2], nb)
    assert 'repulsion724' in sphere.evaluation_hemisphere()[3]
