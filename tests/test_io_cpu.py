"""Host-side file formats of the tracking output path (no GPU): NIfTI round trip, TRK / TCK writers
against their readers, ragged and empty batches, multi-batch streaming, and the Tracker's pass planning."""
import struct

import numpy as np
import pytest

from tracktolearn_b200.io import nifti
from tracktolearn_b200.io.streamlines import TckWriter, TrkWriter, detect_format, read_tck, read_trk


def _ragged(rs, n, lo=1, hi=40):
    lens = rs.randint(lo, hi, size=n)
    data = rs.normal(size=(int(lens.sum()), 3)).astype(np.float32) * 30
    offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    return data, offsets


def test_nifti_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    affine = np.diag([1.25, 1.25, 1.25, 1.0])
    affine[:3, 3] = [-90.0, -126.0, -72.0]
    for arr in (rs.normal(size=(7, 8, 6, 45)).astype(np.float32), (rs.uniform(size=(7, 8, 6)) > 0.5).astype(np.uint8)):
        for ext in ('.nii', '.nii.gz'):
            p = str(tmp_path / ('x' + ext))
            nifti.save(p, arr, affine)
            img = nifti.load(p)
            assert tuple(img.shape) == arr.shape
            np.testing.assert_array_equal(img.get_fdata(dtype=np.float32), arr.astype(np.float32))
            np.testing.assert_allclose(img.affine, affine, atol=1e-5)
            np.testing.assert_allclose(img.zooms[:3], [1.25] * 3, atol=1e-6)


@pytest.mark.parametrize('fmt', ['trk', 'tck'])
def test_writers_round_trip_in_batches(tmp_path, fmt):
    rs = np.random.RandomState(1)
    affine = np.diag([2.0, 2.0, 2.0, 1.0])
    path = str(tmp_path / ('t.' + fmt))
    assert detect_format(path) == fmt
    w = TrkWriter(path, (20, 22, 18), (2.0, 2.0, 2.0), affine, save_seeds=True) if fmt == 'trk' else TckWriter(path)
    all_data, all_lens, all_seeds = [], [], []
    for n in (5, 0, 1, 300):                     # including an empty batch and a single streamline
        data, offsets = _ragged(rs, n)
        seeds = rs.normal(size=(n, 3))
        w.write(data, offsets, seeds)
        all_data.append(data)
        all_lens.append(np.diff(offsets))
        all_seeds.append(seeds)
    w.close()
    assert w.n == 306
    data, offsets, hdr = (read_trk if fmt == 'trk' else read_tck)(path)
    np.testing.assert_array_equal(np.diff(offsets), np.concatenate(all_lens))
    np.testing.assert_array_equal(data, np.concatenate(all_data))
    if fmt == 'trk':
        assert hdr['n_count'] == 306
        with open(path, 'rb') as f:
            raw = f.read()
        assert raw[:5] == b'TRACK' and struct.unpack('<i', raw[996:1000])[0] == 1000


def test_detect_format_rejects_other_extensions():
    with pytest.raises(Exception):
        detect_format('out.vtk')


def test_tracker_pass_planning():
    """Streaming passes cover the seed list exactly once, in order; the buffer budget bounds a pass."""
    from tracktolearn_b200.tracking.tracker import Tracker

    class Env(object):
        max_nb_steps = 798
        seeds = np.zeros((1000003, 3))

    t = Tracker(alg=None, n_actor=50000)
    passes = list(t._passes(Env()))
    assert passes[0][0] == 0 and passes[-1][1] == 1000003
    assert all(a[1] == b[0] for a, b in zip(passes[:-1], passes[1:]))
    assert all(s == 50000 for _, _, s in passes)
    assert max(e - s for s, e, _ in passes) * 799 * 12 <= 24e9 + 799 * 12
    t2 = Tracker(alg=None, n_actor=4096, streaming=False)
    p2 = list(t2._passes(Env()))
    assert all(e - s <= 4096 and slots is None for s, e, slots in p2) and p2[-1][1] == 1000003
