"""Independent cross-checks of the oracle's restatements of THIRD-PARTY algorithms whose sources are absent
from /root/reference (dwi_ml trilinear interpolation, dipy set_number_of_points / length).  The reference
fixtures pin the oracle as a whole (tests/test_oracle_golden.py, through stubs written beside it); here each
restated piece is compared with an implementation of the same published algorithm that somebody else wrote
and that this image ships: scipy.ndimage.map_coordinates(order=1) and numpy.interp.  This does not replace a
run against upstream dwi_ml / dipy (DESIGN.md section 2 keeps "unpinned" for them); it rules out a shared
misunderstanding between the oracle and the stub."""
import numpy as np
from scipy.ndimage import map_coordinates

from oracle import ttl_oracle as O


def test_trilinear_matches_scipy_linear_interpolation_everywhere():
    rs = np.random.RandomState(0)
    vol = rs.normal(size=(9, 11, 7, 5)).astype(np.float32)
    # interior, on-lattice, on the faces, and outside the volume on every side (corner indices clamp:
    # the same extension as scipy's mode='nearest')
    pts = np.concatenate([
        rs.uniform(0, 1, size=(400, 3)) * (np.asarray(vol.shape[:3]) - 1),
        rs.randint(0, 7, size=(50, 3)).astype(np.float64),
        rs.uniform(-2.5, 12.5, size=(400, 3)),
        np.asarray([[0, 0, 0], [8, 10, 6], [8.0, 3.3, 6.0], [-0.5, 10.5, 3.0], [-1.0, -1.0, -1.0]])]).astype(np.float32)
    got = O.trilinear(vol, pts)
    want = np.stack([map_coordinates(vol[..., c].astype(np.float64), pts.T.astype(np.float64), order=1, mode='nearest')
                     for c in range(vol.shape[3])], 1)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


def test_neighbourhood_interpolation_is_seven_trilinear_lookups():
    rs = np.random.RandomState(1)
    vol = rs.normal(size=(8, 8, 8, 3)).astype(np.float32)
    pts = rs.uniform(1.5, 5.5, size=(20, 3)).astype(np.float32)
    r = 0.6
    nb = O.neighborhood_directions(r)
    got = O.interpolate_in_neighborhood(vol, pts, nb).reshape(20, 7, 3)
    offs = np.asarray([[0, 0, 0], [r, 0, 0], [0, r, 0], [0, 0, r], [-r, 0, 0], [0, -r, 0], [0, 0, -r]], dtype=np.float32)
    for k in range(7):
        q = (pts + offs[k]).astype(np.float64)
        want = np.stack([map_coordinates(vol[..., c].astype(np.float64), q.T, order=1, mode='nearest') for c in range(3)], 1)
        np.testing.assert_allclose(got[:, k], want, rtol=0, atol=2e-6)


def _resample_with_interp(s, n):
    s = np.asarray(s, dtype=np.float64)
    cum = np.concatenate(([0.0], np.cumsum(np.linalg.norm(np.diff(s, axis=0), axis=1))))
    t = np.linspace(0.0, cum[-1], n)
    return np.stack([np.interp(t, cum, s[:, k]) for k in range(3)], 1)


def test_set_number_of_points_is_arc_length_linear_resampling():
    rs = np.random.RandomState(2)
    for n_in, n_out in ((2, 128), (3, 128), (17, 128), (200, 128), (300, 12), (128, 128)):
        s = np.cumsum(rs.normal(size=(n_in, 3)), axis=0).astype(np.float32)
        got = O.set_number_of_points(s, n_out)
        assert got.dtype == np.float32 and got.shape == (n_out, 3)
        want = _resample_with_interp(s, n_out)
        scale = np.abs(s).max() + 1.0
        np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * scale)
        np.testing.assert_array_equal(got[0], s[0])
        np.testing.assert_array_equal(got[-1], s[-1])
        # equal spacing along the ORIGINAL polyline
        seg = np.linalg.norm(np.diff(want, axis=0), axis=1)
        assert seg.max() <= O.streamline_length(s) / (n_out - 1) + 1e-9


def test_streamline_length_and_features():
    rs = np.random.RandomState(3)
    s = np.cumsum(rs.normal(size=(40, 3)), axis=0).astype(np.float32)
    assert abs(O.streamline_length(s) - np.linalg.norm(np.diff(s.astype(np.float64), axis=0), axis=1).sum()) < 1e-9
    f = O.oracle_features([s, s[::-1].copy()], 128)
    assert f.shape == (2, 127, 3) and f.dtype == np.float32
    # resampled segments all have (nearly) the same length, bounded by arc length / 127
    n = np.linalg.norm(f[0].astype(np.float64), axis=1)
    assert n.max() <= O.streamline_length(s) / 127 * (1 + 1e-5)
    # reversing the streamline reverses and negates the features (symmetric algorithm up to float rounding)
    np.testing.assert_allclose(f[1], -f[0][::-1], rtol=0, atol=2e-4)


def test_hemisphere_edges_and_local_maxima_against_a_convex_hull_triangulation():
    """The evaluation sphere's edge table against scipy's Delaunay triangulation of the same points on the
    sphere (the convex hull of a centrally symmetric point set), and the oracle's edge-walking
    ``local_maxima`` (dipy's algorithm restated) against a brute-force neighbourhood scan over those hull edges."""
    from scipy.spatial import ConvexHull
    from tracktolearn_b200.datasets.sphere import hemisphere
    v, e, nb = hemisphere(3)
    full = np.concatenate([v, -v])                                  # antipodal copies: index i + V  <->  i
    V = len(v)
    hull = ConvexHull(full)
    he = np.concatenate([hull.simplices[:, [0, 1]], hull.simplices[:, [1, 2]], hull.simplices[:, [2, 0]]]) % V
    he = np.sort(he[he[:, 0] != he[:, 1]], axis=1)
    he = np.unique(he, axis=0)
    # The icosphere's faces are a valid Delaunay triangulation except where four points are cocircular (hull picks
    # either diagonal): every product edge whose endpoints are closer than the shortest non-edge must be a hull edge.
    mine = {tuple(x) for x in e.tolist()}
    theirs = {tuple(x) for x in he.tolist()}
    assert len(mine) == len(theirs)
    common = mine & theirs
    assert len(common) >= 0.9 * len(mine)
    ang = lambda a, b: np.degrees(np.arccos(np.clip(abs(float(v[a] @ v[b])), -1, 1)))
    longest = max(ang(a, b) for a, b in mine)
    for a, b in theirs - mine:                                      # alternative diagonals only: same length class
        assert ang(a, b) <= longest * 1.25
    # brute force over the hull's neighbourhoods on smooth symmetric functions: strict maxima agree
    rs = np.random.RandomState(5)
    neigh = [[] for _ in range(V)]
    for a, b in e:
        neigh[a].append(b)
        neigh[b].append(a)
    for _ in range(20):
        axes = np.linalg.qr(rs.normal(size=(3, 3)))[0].T               # three orthogonal lobes: none swallows another
        w = rs.uniform(0.3, 1.0, size=3)
        odf = sum(wk * np.abs(v @ ax) ** 8 for wk, ax in zip(w, axes))
        vals, idx = O.local_maxima(odf, e)
        brute = [i for i in range(V) if all(odf[i] >= odf[j] for j in neigh[i]) and any(odf[i] > odf[j] for j in neigh[i])]
        assert sorted(idx.tolist()) == sorted(brute)
        assert np.all(np.diff(vals) <= 0)
        # every lobe axis has a detected maximum within the sphere's resolution (8 degrees between vertices)
        for ax in axes:
            assert max(abs(float(v[i] @ ax)) for i in idx) > np.cos(np.deg2rad(9.0))
