"""Stand-ins for the third-party packages the reference imports but this container lacks.

Used ONLY by make_golden.py, in the build container, to import the reference's own
Python (``/root/reference/TrackToLearn``) and record what ITS code computes.  Nothing
here is imported by the product, the tests or the bench at run time.

What is real and what is restated:
  * everything under ``TrackToLearn.*`` is the unmodified reference;
  * ``scipy.ndimage`` (mask criterion), ``torch`` and ``numpy`` are the real packages;
  * ``dwi_ml`` (pinned by the reference only to a branch, requirements.txt:1) is
    absent: ``interpolate_volume_in_neighborhood`` / ``get_neighborhood_vectors_axes``
    below restate its published algorithm -- for that one
    function the fixtures pin the restatement, not upstream dwi_ml ("parity
    unpinned" for it, stated in DESIGN.md);
  * ``dipy`` / ``nibabel`` / ``scilpy`` / ``h5py`` are absent: the handful of
    callables the hot path touches (``random_seeds_from_mask``, ``Tractogram``,
    ``StatefulTractogram``, ``set_number_of_points``, ``length``) are restated, the
    rest are inert placeholders so that module-level imports succeed.
"""
import sys
import types

import numpy as np
import torch


class _Inert(types.ModuleType):
    """Module whose every missing attribute is an inert callable placeholder."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)

        class _Placeholder:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                raise RuntimeError('placeholder %s.%s was called' % (self.__class__.__module__, name))
        _Placeholder.__name__ = name
        setattr(self, name, _Placeholder)
        return _Placeholder


def _mod(name):
    m = _Inert(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent:
        if parent not in sys.modules:
            _mod(parent)
        setattr(sys.modules[parent], child, m)
    return m


# ----------------------------------------------------------------------------- dwi_ml
_B1 = torch.tensor([[1, 0, 0, 0, 0, 0, 0, 0],
                    [-1, 0, 0, 0, 1, 0, 0, 0],
                    [-1, 0, 1, 0, 0, 0, 0, 0],
                    [-1, 1, 0, 0, 0, 0, 0, 0],
                    [1, 0, -1, 0, -1, 0, 1, 0],
                    [1, -1, -1, 1, 0, 0, 0, 0],
                    [1, -1, 0, 0, -1, 1, 0, 0],
                    [-1, 1, 1, -1, 1, -1, -1, 1]], dtype=torch.float32)
_IDX_BOX = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                         [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=torch.float32)


def torch_trilinear_interpolation(volume, coords_vox_corner):
    device = volume.device
    B1 = _B1.to(device)
    idx_box = _IDX_BOX.to(device)
    if volume.dim() == 3:
        volume = volume.unsqueeze(-1)
    idx = torch.floor(coords_vox_corner[:, None, :] + idx_box[None]).reshape((-1, 3)).long()
    lower = torch.zeros(3, dtype=torch.long, device=device)
    upper = torch.as_tensor(volume.shape[:3], device=device) - 1
    idx = torch.min(torch.max(idx, lower), upper)
    d = coords_vox_corner - torch.floor(coords_vox_corner)
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    Q1 = torch.stack([torch.ones_like(dx), dx, dy, dz, dx * dy, dy * dz, dx * dz,
                      dx * dy * dz], dim=0)
    P = volume[idx[:, 0], idx[:, 1], idx[:, 2]]
    P = P.reshape((coords_vox_corner.shape[0], 8, volume.shape[-1]))
    return torch.sum(P * torch.mm(Q1.t(), B1)[:, :, None], dim=1)


def interpolate_volume_in_neighborhood(volume_as_tensor, coords_vox_corner,
                                       neighborhood_vectors_vox=None, clear_cache=True):
    if neighborhood_vectors_vox is not None:
        m = coords_vox_corner.shape[0]
        n = neighborhood_vectors_vox.shape[0]
        coords = coords_vox_corner[:, None, :] + neighborhood_vectors_vox[None, :, :]
        coords = coords.reshape((m * n, 3))
        out = torch_trilinear_interpolation(volume_as_tensor, coords)
        out = out.reshape((m, -1))
        return out, coords
    return torch_trilinear_interpolation(volume_as_tensor, coords_vox_corner), coords_vox_corner


def get_neighborhood_vectors_axes(radius, resolution):
    # dwi_ml: unit axes (+x,+y,+z,-x,-y,-z), one shell per radius step
    axes = torch.eye(3)
    unit = torch.cat((axes, -axes))
    out = []
    for r in range(int(radius)):
        out.append(unit * ((r + 1) * resolution))
    return torch.cat(out)


# ------------------------------------------------------------------------------- dipy
def random_seeds_from_mask(mask, affine, seeds_count=1, seed_count_per_voxel=True,
                           random_seed=None):
    mask = np.array(mask, dtype=bool, ndmin=3)
    where = np.argwhere(mask)
    seeds = []
    for i in range(1, seeds_count + 1):
        for s in where:
            grid = np.random.random(3)
            seeds.append(s + grid - .5)
    seeds = np.asarray(seeds)
    if seeds.any():
        seeds = seeds @ affine[:3, :3].T + affine[:3, 3]
    return seeds


def length(streamline):
    s = np.asarray(streamline, dtype=np.float64)
    if len(s) < 2:
        return 0.0
    return float(np.sqrt(((s[1:] - s[:-1]) ** 2).sum(-1)).sum())


def set_number_of_points(streamlines, nb_points=3):
    """dipy c_set_number_of_points restated (Cython fused-type semantics: segment
    differences are taken in the input dtype, arc lengths and the interpolation in
    double, the result is stored in the input dtype; the last point is copied)."""
    single = isinstance(streamlines, np.ndarray) and streamlines.ndim == 2
    if single:
        streamlines = [streamlines]
    out = []
    for s in streamlines:
        s = np.asarray(s)
        N = len(s)
        d = np.diff(s, axis=0).astype(np.float64)
        seg = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        cum = np.zeros(N, dtype=np.float64)
        for i in range(1, N):
            cum[i] = cum[i - 1] + seg[i - 1]
        step = cum[N - 1] / (nb_points - 1)
        res = np.zeros((nb_points, 3), dtype=s.dtype)
        nxt = 0.0
        i = 0
        k = 0
        while nxt < cum[N - 1]:
            if nxt == cum[k]:
                res[i] = s[k]
                nxt += step
                i += 1
                k += 1
            elif nxt < cum[k]:
                ratio = 1 - ((cum[k] - nxt) / (cum[k] - cum[k - 1]))
                delta = (s[k] - s[k - 1]).astype(np.float64)
                res[i] = s[k - 1].astype(np.float64) + ratio * delta
                nxt += step
                i += 1
            else:
                k += 1
            if i >= nb_points:
                break
        res[nb_points - 1] = s[N - 1]
        out.append(res)
    return out[0] if single else out


class Tractogram:
    def __init__(self, streamlines=None, data_per_streamline=None, data_per_point=None,
                 affine_to_rasmm=None):
        self.streamlines = [np.array(s, copy=True) for s in (streamlines if streamlines is not None else [])]
        self.data_per_streamline = data_per_streamline or {}

    def apply_affine(self, affine, lazy=False):
        # nibabel's ArraySequence keeps float32 storage
        self.streamlines = [(s @ affine[:3, :3].T + affine[:3, 3]).astype(np.float32) for s in self.streamlines]
        return self

    def __len__(self):
        return len(self.streamlines)


class _Space:
    RASMM = 'rasmm'
    VOX = 'vox'


class StatefulTractogram:
    """Just enough for oracle_reward / OracleStoppingCriterion: RASMM -> vox/corner."""

    def __init__(self, streamlines, reference, space, origin=None, **k):
        self._sl = [np.asarray(s) for s in streamlines]
        self._inv = np.linalg.inv(np.asarray(reference))
        self.space = space

    def to_vox(self):
        self._sl = [(s @ self._inv[:3, :3].T + self._inv[:3, 3]).astype(np.float32) for s in self._sl]

    def to_corner(self):
        self._sl = [(s + np.float32(0.5)).astype(np.float32) for s in self._sl]

    @property
    def streamlines(self):
        return self._sl


def install():
    for name in ['nibabel', 'nibabel.streamlines', 'nibabel.streamlines.tractogram',
                 'dipy', 'dipy.core', 'dipy.core.sphere', 'dipy.core.geometry', 'dipy.data',
                 'dipy.direction', 'dipy.direction.peaks', 'dipy.tracking',
                 'dipy.tracking.utils', 'dipy.tracking.metrics', 'dipy.tracking.streamline',
                 'dipy.tracking.streamlinespeed', 'dipy.reconst', 'dipy.reconst.shm',
                 'dipy.reconst.csdeconv', 'dipy.io', 'dipy.io.stateful_tractogram',
                 'dipy.io.utils', 'dwi_ml', 'dwi_ml.data', 'dwi_ml.data.processing',
                 'dwi_ml.data.processing.volume', 'dwi_ml.data.processing.volume.interpolation',
                 'dwi_ml.data.processing.space', 'dwi_ml.data.processing.space.neighborhood',
                 'scilpy', 'scilpy.reconst', 'scilpy.reconst.utils', 'scilpy.reconst.sh',
                 'scilpy.io', 'scilpy.io.utils', 'scilpy.tracking', 'scilpy.tracking.utils',
                 'h5py', 'comet_ml']:
        _mod(name)
    sm = sys.modules
    sm['dwi_ml.data.processing.volume.interpolation'].interpolate_volume_in_neighborhood = \
        interpolate_volume_in_neighborhood
    sm['dwi_ml.data.processing.volume.interpolation'].torch_trilinear_interpolation = \
        torch_trilinear_interpolation
    sm['dwi_ml.data.processing.space.neighborhood'].get_neighborhood_vectors_axes = \
        get_neighborhood_vectors_axes
    sm['dipy.tracking.utils'].random_seeds_from_mask = random_seeds_from_mask
    sm['dipy.tracking.streamline'].set_number_of_points = set_number_of_points
    sm['dipy.tracking.streamlinespeed'].length = length
    sm['dipy.tracking.streamlinespeed'].set_number_of_points = set_number_of_points
    sm['nibabel.streamlines'].Tractogram = Tractogram
    sm['dipy.io.stateful_tractogram'].Tractogram = Tractogram
    sm['dipy.io.stateful_tractogram'].StatefulTractogram = StatefulTractogram
    sm['dipy.io.stateful_tractogram'].Space = _Space
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
