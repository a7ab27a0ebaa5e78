"""File loading for tracking (reference: environments/env.py:350-449 ``_load_files`` and
datasets/utils.py:127-179 ``set_sh_order_basis``), without nibabel / dipy / scilpy."""
import numpy as np

from tracktolearn_b200.datasets.utils import MRIDataVolume, get_sh_order_and_fullness
from tracktolearn_b200.io import nifti


def sph_harm_ind_list(sh_order, full_basis=False):
    """dipy ``sph_harm_ind_list``: (m_list, l_list) of the real SH basis."""
    ls = range(0, sh_order + 1, 1 if full_basis else 2)
    m_list, l_list = [], []
    for l in ls:
        for m in range(-l, l + 1):
            m_list.append(m)
            l_list.append(l)
    return np.asarray(m_list), np.asarray(l_list)


def convert_sh_basis_legacy(sh, order):
    """tournier07 <-> descoteaux07 (legacy conventions).  The two real bases use
    sqrt(2)Re(Y_l^|m|) and sqrt(2)Im(Y_l^|m|) on opposite signs of m, so the change of basis is
    the permutation m <-> -m inside every degree l (an involution).  The reference reaches the
    same coefficients numerically through scilpy ``convert_sh_basis`` (SH -> SF on repulsion724
    -> SH least squares, datasets/utils.py:172-177)."""
    m_list, l_list = sph_harm_ind_list(order)
    perm = np.empty(len(m_list), dtype=np.int64)
    index = {(int(l), int(m)): i for i, (m, l) in enumerate(zip(m_list, l_list))}
    for i, (m, l) in enumerate(zip(m_list, l_list)):
        perm[i] = index[(int(l), int(-m))]
    return sh[..., perm]


def set_sh_order_basis(sh, sh_basis, target_basis='descoteaux07', target_order=6):
    """Reference: datasets/utils.py:127-179."""
    n_coefs = sh.shape[-1]
    sh_order, full_basis = get_sh_order_and_fullness(n_coefs)
    sh_order = int(sh_order)
    target_order = int(target_order)
    if full_basis:
        print('SH coefficients are in "full" basis, only even coefficients will be used.')
        _, orders = sph_harm_ind_list(sh_order, True)
        sh = sh[..., orders % 2 == 0]
        n_coefs = sh.shape[-1]
    if sh_order != target_order:
        print('SH coefficients are of order {}, converting them to order {}.'.format(sh_order, target_order))
        target_n_coefs = len(sph_harm_ind_list(target_order)[0])
        if n_coefs > target_n_coefs:
            sh = sh[..., :target_n_coefs]
        else:
            X, Y, Z = sh.shape[:3]
            sh = np.concatenate((sh, np.zeros((X, Y, Z, target_n_coefs - n_coefs), dtype=sh.dtype)), axis=-1)
    if sh_basis != target_basis:
        print('SH coefficients are in the {} basis, converting them to {}.'.format(sh_basis, target_basis))
        sh = convert_sh_basis_legacy(sh, target_order)
    return np.ascontiguousarray(sh, dtype=np.float32)


def load_files(signal_file, in_seed, in_mask, sh_basis, target_sh_order=6, compute_peaks=False,
               device='cuda:0'):
    """Reference: environments/env.py:350-449.  Peaks are only needed for the alignment reward,
    which ``ttl_track`` never computes (compute_reward=False, ttl_track.py:80); the reference extracts
    them for every voxel in a Python loop regardless (env.py:417-425, minutes on a whole brain).  Here
    they are computed only when asked for, by one kernel launch (datasets/peaks.py), on this package's
    own 321-direction hemisphere instead of dipy's repulsion724 (absent offline)."""
    signal = nifti.load(signal_file)
    if not np.allclose(np.mean(signal.zooms[:3]), signal.zooms[0], atol=1e-03):
        print('WARNING: ODF SH file is not isotropic. Tracking cannot be ran robustly. You are '
              'entering undefined behavior territory.')
    data = set_sh_order_basis(signal.get_fdata(dtype=np.float32), sh_basis,
                              target_order=target_sh_order, target_basis='descoteaux07')
    peaks_volume = None
    if compute_peaks:
        from tracktolearn_b200.datasets.peaks import compute_peaks as _compute_peaks
        peaks_volume = MRIDataVolume(_compute_peaks(data, device=device).cpu().numpy(), signal.affine)
    seeding = nifti.load(in_seed)
    tracking = nifti.load(in_mask)
    signal_volume = MRIDataVolume(data, signal.affine)
    seeding_volume = MRIDataVolume(seeding.get_fdata(), seeding.affine)
    tracking_volume = MRIDataVolume(tracking.get_fdata(), tracking.affine)
    return (signal_volume, peaks_volume, tracking_volume, seeding_volume)
