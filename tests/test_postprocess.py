"""Tractogram post-processing: dipy `length` and `compress_streamlines` on
packed arrays.  CPU: properties of the oracle restatement; GPU: the kernels against the oracle."""
import numpy as np
import pytest

from oracle import ttl_oracle as O


def _curves(rs, n, dtype=np.float32):
    out = []
    for i in range(n):
        N = int(rs.randint(1, 140))
        t = np.cumsum(rs.uniform(0.2, 1.0, size=N))
        k = rs.uniform(0.02, 0.3)
        s = np.stack([t * rs.uniform(0.5, 1.5), np.sin(k * t) * rs.uniform(1, 20), np.cos(0.5 * k * t) * 5], 1)
        s += rs.normal(0, rs.choice([0.0, 1e-3, 0.05]), size=s.shape)
        out.append(s.astype(dtype))
    # edge cases: 1-3 points, exactly collinear, repeated points (zero chord -> NaN distance), long segments
    out.append(np.zeros((1, 3), dtype))
    out.append(np.asarray([[0, 0, 0], [1, 2, 3]], dtype))
    out.append(np.asarray([[0, 0, 0], [1, 1, 1], [2, 2, 2]], dtype))
    out.append(np.outer(np.arange(50), [1.0, 0.5, 0.25]).astype(dtype))
    out.append(np.outer(np.arange(50), [9.0, 0.0, 0.0]).astype(dtype))
    rep = np.outer(np.arange(20), [0.3, 0.1, 0.0]).astype(dtype)
    rep[5:9] = rep[5]
    out.append(rep)
    loop = np.stack([np.cos(np.linspace(0, 2 * np.pi, 40)), np.sin(np.linspace(0, 2 * np.pi, 40)), np.zeros(40)], 1)
    loop[-1] = loop[0]
    out.append((3 * loop).astype(dtype))
    return out


def _pack(sl):
    offsets = np.concatenate(([0], np.cumsum([len(s) for s in sl]))).astype(np.int64)
    return np.concatenate(sl).astype(np.float32), offsets


def test_compress_oracle_properties():
    rs = np.random.RandomState(5)
    for s in _curves(rs, 60):
        for tol in (0.001, 0.01, 0.2):
            c = O.compress_streamline(s, tol)
            assert c.dtype == s.dtype
            if len(s) <= 2:
                np.testing.assert_array_equal(c, s)
                continue
            # end points kept, points are a subsequence in order, no chord longer than 10 unless it is
            # an original segment
            np.testing.assert_array_equal(c[0], s[0])
            np.testing.assert_array_equal(c[-1], s[-1])
            j = 0
            idx = []
            for p in c:
                while not np.array_equal(s[j], p):
                    j += 1
                idx.append(j)
                j += 1 if len(idx) < len(c) else 0
            assert idx == sorted(idx)
            # every dropped point is within tol of its chord's line
            for a, b in zip(idx[:-1], idx[1:]):
                if b - a < 2:
                    continue
                u = (s[b].astype(np.float64) - s[a])
                for k in range(a + 1, b):
                    w = s[k].astype(np.float64) - s[b]
                    d = np.linalg.norm(np.cross(u, w)) / np.linalg.norm(u)
                    assert d <= tol * (1 + 1e-4) + 1e-6, (d, tol)
    # a straight line within the segment-length bound collapses to its end points
    line = np.outer(np.linspace(0, 9, 30), [1.0, 0.0, 0.0]).astype(np.float32)
    assert len(O.compress_streamline(line, 0.01)) == 2
    # ... and is cut into chords shorter than max_segment_length beyond it
    long_line = np.outer(np.linspace(0, 95, 200), [1.0, 0.0, 0.0]).astype(np.float32)
    c = O.compress_streamline(long_line, 0.01)
    assert 10 <= len(c) <= 12 and np.all(np.linalg.norm(np.diff(c, axis=0), axis=1) < 10.0)
    # idempotent on its own output for a tolerance well above float rounding
    rs = np.random.RandomState(6)
    for s in _curves(rs, 10):
        c = O.compress_streamline(s, 0.05)
        assert len(O.compress_streamline(c, 0.05)) <= len(c)


@pytest.mark.gpu
def test_compress_and_length_kernels_match_oracle():
    import torch
    from tracktolearn_b200.tracking.postprocess import compress_packed, lengths_packed
    rs = np.random.RandomState(9)
    sl = _curves(rs, 400)
    data, offsets = _pack(sl)
    lens = lengths_packed(data, offsets, device='cuda:0').cpu().numpy()
    ref = np.asarray([O.streamline_length(s) for s in sl])
    np.testing.assert_allclose(lens, ref, rtol=1e-12, atol=1e-12)
    for tol in (0.001, 0.01, 0.2):
        d, o = compress_packed(data, offsets, tol_error=tol, device='cuda:0')
        d, o = d.cpu().numpy(), o.cpu().numpy()
        assert o[0] == 0 and o[-1] == len(d)
        for i, s in enumerate(sl):
            want = O.compress_streamline(s, tol)
            got = d[o[i]:o[i + 1]]
            assert got.shape == want.shape, (i, len(s), got.shape, want.shape)
            np.testing.assert_array_equal(got, want)     # bit-exact: same points selected
    # empty input
    d, o = compress_packed(np.zeros((0, 3), np.float32), np.zeros((1,), np.int64), device='cuda:0')
    assert d.shape[0] == 0 and o.tolist() == [0]
    assert torch.cuda.is_available()


@pytest.mark.gpu
def test_tracker_compress_option():
    """Tracker(compress=t): what `ttl_track.py --compress t` yields -- every kept streamline is the
    oracle's compression of the uncompressed one."""
    import torch
    from tests.test_tracker_gpu import _setup
    from tracktolearn_b200.tracking.tracker import Tracker
    env, alg, sub, seeds, sd = _setup(precision='bf16')
    seeds0 = np.array(seeds)
    env.seeds = seeds0.copy()
    np.random.seed(0)
    plain = list(Tracker(alg, 256, min_length=5, max_length=200).track(env, 'tck'))
    env.seeds = seeds0.copy()
    np.random.seed(0)
    comp = list(Tracker(alg, 256, compress=0.05, min_length=5, max_length=200).track(env, 'tck'))
    assert len(plain) == len(comp) > 50
    shorter = 0
    for a, b in zip(plain, comp):
        want = O.compress_streamline(a.streamline.astype(np.float32), 0.05)
        np.testing.assert_allclose(b.streamline, want, rtol=0, atol=0)
        shorter += len(b.streamline) < len(a.streamline)
    assert shorter > len(plain) // 2
    assert torch.cuda.is_available()


@pytest.mark.gpu
def test_tracker_compress_tolerance_is_in_mm():
    """`--compress` is given in mm; the streamlines are compressed in voxel space with
    compress / voxel_size (reference tracking/tracker.py:103,123-125).  On a 2 mm volume a 0.05 mm
    tolerance is 0.025 voxels."""
    from tests.test_tracker_gpu import _setup
    from tracktolearn_b200.tracking.tracker import Tracker
    env, alg, sub, seeds, sd = _setup(precision='fp16', vox=2.0)
    assert abs(np.mean(np.abs(np.diag(env.affine_vox2rasmm))[:3]) - 2.0) < 1e-12
    seeds0 = np.array(seeds)
    env.seeds = seeds0.copy()
    np.random.seed(0)
    plain = list(Tracker(alg, 256, min_length=5, max_length=200).track(env, 'tck'))
    env.seeds = seeds0.copy()
    np.random.seed(0)
    comp = list(Tracker(alg, 256, compress=0.05, min_length=5, max_length=200).track(env, 'tck'))
    assert len(plain) == len(comp) > 50
    differs_from_unscaled = 0
    for a, b in zip(plain, comp):
        vox = (a.streamline / 2.0).astype(np.float32)            # tck space = 2 * voxel space, exactly
        want = O.compress_streamline(vox, 0.05 / 2.0).astype(np.float64) * 2.0
        np.testing.assert_allclose(b.streamline, want, rtol=0, atol=1e-6)
        differs_from_unscaled += len(O.compress_streamline(vox, 0.05)) != len(want)
    assert differs_from_unscaled > 0        # the unscaled tolerance would have kept fewer points
