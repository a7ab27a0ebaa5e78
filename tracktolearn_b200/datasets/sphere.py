"""Evaluation sphere for load-time peak extraction (reference: environments/env.py:411-414 uses
``HemiSphere.from_sphere(get_sphere("repulsion724")).subdivide(0)`` -- 362 directions with the edges of
their triangulation).  dipy and its sphere data file are not available offline, so the product builds
its own centrally symmetric sphere: an icosahedron subdivided three times (642 vertices, 8 degrees
apart), one representative per antipodal pair (321 directions), edges of the full triangulation mapped
onto the representatives -- the same construction as dipy's HemiSphere, a different point set.  Peaks
extracted on it agree with the reference's to the angular resolution of the two spheres (~4 degrees);
see DESIGN.md."""
import numpy as np


def icosphere(subdivisions=3):
    """-> (vertices [V,3] float64 on the unit sphere, faces [F,3] int64)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.asarray([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                    [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.asarray([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                    [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                    [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdivisions):
        verts = [tuple(p) for p in v]
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.asarray(verts[a]) + np.asarray(verts[b])) / 2.0
                m /= np.linalg.norm(m)
                cache[key] = len(verts)
                verts.append(tuple(m))
            return cache[key]
        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.asarray(verts, dtype=np.float64), np.asarray(nf, dtype=np.int64)
    return v, f


def hemisphere(subdivisions=3):
    """-> (vertices [V,3] float64, edges [E,2] int64 with a < b, neighbours [V,D] int32 padded with -1).
    One representative per antipodal pair of the icosphere; an edge (a, b) of the full sphere becomes
    (rep(a), rep(b)), so a symmetric function's local maxima are found across the equator too."""
    v, f = icosphere(subdivisions)
    key = {tuple(np.round(p, 9) + 0.0): i for i, p in enumerate(v)}
    anti = np.asarray([key[tuple(np.round(-p, 9) + 0.0)] for p in v])
    eps = 1e-12
    upper = (v[:, 2] > eps) | ((np.abs(v[:, 2]) <= eps) & ((v[:, 1] > eps) | ((np.abs(v[:, 1]) <= eps) & (v[:, 0] > 0))))
    assert (upper != upper[anti]).all()
    keep = np.nonzero(upper)[0]
    new_index = -np.ones(len(v), dtype=np.int64)
    new_index[keep] = np.arange(len(keep))
    rep = np.where(upper, new_index, new_index[anti])
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e = rep[e]
    e = np.sort(e[e[:, 0] != e[:, 1]], axis=1)
    e = np.unique(e, axis=0)
    V = len(keep)
    nb = [[] for _ in range(V)]
    for a, b in e:
        nb[a].append(int(b))
        nb[b].append(int(a))
    D = max(len(x) for x in nb)
    table = -np.ones((V, D), dtype=np.int32)
    for i, x in enumerate(nb):
        table[i, :len(x)] = sorted(x)
    return np.ascontiguousarray(v[keep]), e, table


def _tables_from_edges(vertices, edges):
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    e = np.sort(np.asarray(edges, dtype=np.int64).reshape(-1, 2), axis=1)
    e = np.unique(e[e[:, 0] != e[:, 1]], axis=0)
    V = len(vertices)
    nb = [[] for _ in range(V)]
    for a, b in e:
        nb[a].append(int(b))
        nb[b].append(int(a))
    D = max(len(x) for x in nb)
    table = -np.ones((V, D), dtype=np.int32)
    for i, x in enumerate(nb):
        table[i, :len(x)] = sorted(x)
    return vertices, e, table


def reference_hemisphere():
    """The sphere the reference evaluates peaks on (environments/env.py:412-415):
    ``HemiSphere.from_sphere(get_sphere("repulsion724")).subdivide(0)`` -- 362 directions and the edges of
    their triangulation -- taken from a dipy installation when there is one, so that load-time peaks are
    comparable one to one with the reference's.  Returns (vertices, edges, neighbours) like ``hemisphere``,
    or None when dipy (or its data file) is not importable."""
    try:
        try:
            from dipy.data import get_sphere          # dipy >= 1.7: get_sphere(name=...)
        except ImportError:
            return None
        from dipy.core.sphere import HemiSphere
        try:
            full = get_sphere(name='repulsion724')
        except TypeError:
            full = get_sphere('repulsion724')
        hemi = HemiSphere.from_sphere(full).subdivide(0)
        return _tables_from_edges(np.asarray(hemi.vertices), np.asarray(hemi.edges))
    except Exception:
        return None


def evaluation_hemisphere():
    """(vertices, edges, neighbours, name): dipy's repulsion724 hemisphere when available (the reference's),
    else this package's icosphere hemisphere."""
    ref = reference_hemisphere()
    if ref is not None:
        return ref + ('dipy repulsion724 hemisphere (%d directions)' % len(ref[0]),)
    v, e, nb = hemisphere(3)
    return v, e, nb, 'icosphere(3) hemisphere (%d directions; dipy not importable)' % len(v)
