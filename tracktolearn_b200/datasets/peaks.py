"""Load-time peak extraction on the device (reference: environments/env.py:405-432).

``compute_peaks(sh)`` turns an order-N descoteaux07 (legacy) SH volume into the [X, Y, Z, 15] peaks
volume the alignment reward reads (local_reward.py:23-27): up to five fODF maxima per voxel, each scaled
by value / largest value."""
import functools
import math

import numpy as np
import torch

from tracktolearn_b200 import _lib
from tracktolearn_b200.datasets.sphere import evaluation_hemisphere
from tracktolearn_b200.datasets.utils import get_sh_order_and_fullness


def _legendre(lmax, x):
    P = {(0, 0): np.ones_like(x)}
    somx2 = np.sqrt(np.clip(1.0 - x * x, 0.0, None))
    for m in range(1, lmax + 1):
        P[(m, m)] = -(2 * m - 1) * somx2 * P[(m - 1, m - 1)]
    for m in range(0, lmax):
        P[(m + 1, m)] = (2 * m + 1) * x * P[(m, m)]
    for m in range(0, lmax + 1):
        for l in range(m + 2, lmax + 1):
            P[(l, m)] = ((2 * l - 1) * x * P[(l - 1, m)] - (l + m - 1) * P[(l - 2, m)]) / (l - m)
    return P


def sh_to_sf_matrix(vertices, order):
    """Real even SH basis 'descoteaux07' (legacy) on unit vectors -> B [V, n_coefs] float64
    (dipy ``sh_to_sf_matrix(sphere, order, "descoteaux07")``, env.py:414): sqrt(2) Re Y_l^|m| for m < 0,
    Y_l^0, sqrt(2) Im Y_l^m for m > 0, Condon-Shortley phase included."""
    v = np.asarray(vertices, dtype=np.float64)
    ct = np.clip(v[:, 2], -1.0, 1.0)
    phi = np.arctan2(v[:, 1], v[:, 0])
    P = _legendre(order, ct)
    cols = []
    for l in range(0, order + 1, 2):
        for m in range(-l, l + 1):
            am = abs(m)
            norm = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - am) / math.factorial(l + am))
            if m < 0:
                cols.append(math.sqrt(2.0) * norm * P[(l, am)] * np.cos(am * phi))
            elif m == 0:
                cols.append(norm * P[(l, 0)])
            else:
                cols.append(math.sqrt(2.0) * norm * P[(l, am)] * np.sin(am * phi))
    return np.ascontiguousarray(np.stack(cols, axis=1))


@functools.lru_cache(maxsize=4)
def _sphere_tables(order):
    vertices, edges, neighbours, _ = evaluation_hemisphere()     # dipy's repulsion724 when dipy is installed
    return vertices, edges, neighbours, sh_to_sf_matrix(vertices, order)


def compute_peaks(sh, device='cuda:0', npeaks=5, relative_threshold=0.1, absolute_threshold=0.0,
                  min_separation_angle=25.0):
    """sh: [X, Y, Z, C] float32 (numpy or tensor), even descoteaux07 coefficients -> CUDA float32 tensor
    [X, Y, Z, npeaks * 3].  Thresholds as in env.py:419-421 (``get_maximas(..., 0.1, 0)``) and dipy's
    default 25 degree separation."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise _lib.TTLError('compute_peaks runs on a CUDA device only (no CPU fallback)')
    lib = _lib.load()
    t = sh if isinstance(sh, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(sh, dtype=np.float32))
    t = t.to(device, dtype=torch.float32).contiguous()
    C = int(t.shape[-1])
    order, full = get_sh_order_and_fullness(C)
    if full:
        raise ValueError('compute_peaks expects the even (symmetric) SH coefficients')
    vertices, _, neighbours, B = _sphere_tables(int(order))
    dB = torch.from_numpy(B).to(device)
    dV = torch.from_numpy(vertices).to(device)
    dN = torch.from_numpy(neighbours).to(device)
    n_vox = int(np.prod(t.shape[:-1]))
    out = torch.zeros(tuple(t.shape[:-1]) + (npeaks * 3,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.ttl_peaks_from_sh(_lib.ptr(t), n_vox, C, C, _lib.ptr(dB), _lib.ptr(dV), _lib.ptr(dN),
                                         int(vertices.shape[0]), int(neighbours.shape[1]),
                                         float(relative_threshold), float(absolute_threshold),
                                         float(min_separation_angle), int(npeaks), _lib.ptr(out),
                                         _lib.stream_ptr(device)), 'ttl_peaks_from_sh')
    return out
