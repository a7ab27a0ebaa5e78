"""CPU oracle for the TrackToLearn tracking hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain numpy / scipy / torch-CPU restatement of the reference algorithm for the one hot
path this repository accelerates.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it; the product package ``tracktolearn_b200`` never does and fails loudly when
its CUDA library is missing.

Pinning: the reference's tests hold no golden vectors and its third-party
stack cannot be installed here (F9).  The oracle is pinned instead against fixtures
produced by running the reference's OWN code (``/root/reference/TrackToLearn``) in the build
container -- ``tests/golden/make_golden.py``, checked by ``tests/test_oracle_golden.py``.
Two third-party pieces are restated from their published algorithms because their sources
are absent: dwi_ml's trilinear neighbourhood interpolation (pinned by the reference only to
the branch ``for_beluga_scilpy2``, requirements.txt:1) and dipy's ``set_number_of_points`` /
``random_seeds_from_mask`` (unpinned).  For those functions parity is UNPINNED against
upstream: the fixtures pin this restatement.  The same holds for the two
"next-row" pieces added later, both dipy / scilpy algorithms that the reference only calls:
``compress_streamline`` (dipy ``compress_streamlines``, tracker.py:123-125) and
``local_maxima`` / ``peak_directions`` / ``peaks_from_sh`` (scilpy ``get_maximas`` -> dipy
``peak_directions``, env.py:405-432; evaluated on this repository's own sphere because dipy's
``repulsion724`` data file is absent).  They have no reference-recorded fixtures: PARITY UNPINNED.
What narrows the gap: ``tests/test_oracle_independent_cpu.py`` compares ``trilinear`` with
``scipy.ndimage.map_coordinates(order=1, mode='nearest')`` (inside, on and outside the volume) and
``set_number_of_points`` / ``streamline_length`` with ``numpy.interp`` over arc length -- other people's
implementations of the same published algorithms, shipped in this image.

Each function cites the reference lines it follows (paths relative to
``/root/reference/TrackToLearn``).
"""
import math

import numpy as np
from scipy.ndimage import map_coordinates, spline_filter

MASK, LENGTH, CURVATURE, TARGET, LOOP, ANGULAR_ERROR, ORACLE = 1, 2, 4, 8, 16, 32, 64
"""environments/stopping_criteria.py:10-20 (StoppingFlags)."""


def is_flag_set(flags, ref_flag):
    """environments/stopping_criteria.py:23-28."""
    return ((np.asarray(flags).astype(np.uint8) & ref_flag) >>
            np.log2(ref_flag).astype(np.uint8)).astype(bool)


def normalize_vectors(v, norm=1.):
    """utils/utils.py:117-121."""
    return (v / np.sqrt(np.einsum('...i,...i', v, v))[..., None]) * norm


# --------------------------------------------------------------------------------------
# S14: state = trilinear SH at the 7-point neighbourhood + previous directions
# --------------------------------------------------------------------------------------
_B1 = np.array([[1, 0, 0, 0, 0, 0, 0, 0],
                [-1, 0, 0, 0, 1, 0, 0, 0],
                [-1, 0, 1, 0, 0, 0, 0, 0],
                [-1, 1, 0, 0, 0, 0, 0, 0],
                [1, 0, -1, 0, -1, 0, 1, 0],
                [1, -1, -1, 1, 0, 0, 0, 0],
                [1, -1, 0, 0, -1, 1, 0, 0],
                [-1, 1, 1, -1, 1, -1, -1, 1]], dtype=np.float32)
_IDX_BOX = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                     [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=np.float32)


def neighborhood_directions(radius_vox):
    """environments/env.py:207-213: zero vector + dwi_ml get_neighborhood_vectors_axes(1, r)
    = [0; +x; +y; +z; -x; -y; -z] * r, float32."""
    eye = np.eye(3, dtype=np.float32)
    axes = np.concatenate((eye, -eye)) * np.float32(radius_vox)
    return np.concatenate((np.zeros((1, 3), dtype=np.float32), axes)).astype(np.float32)


def trilinear(volume, coords):
    """dwi_ml ``torch_trilinear_interpolation`` restated.

    volume [X,Y,Z,C] float32, coords [n,3] float32 (lattice at integer coordinates).
    Corner indices ``floor(coord + box)`` are clamped to [0, shape-1] independently per
    corner; weights come from the unclamped fractional part through the 8x8 polynomial
    form; everything float32.  NaN coordinates give index 0 and NaN output.
    """
    coords = np.asarray(coords, dtype=np.float32)
    with np.errstate(invalid='ignore'):
        fl = np.floor(coords[:, None, :] + _IDX_BOX[None])
        fl = np.where(np.isnan(fl), -np.inf, fl)          # floor(NaN).long() -> INT64_MIN
        upper = np.asarray(volume.shape[:3], dtype=np.float64) - 1
        idx = np.clip(fl.astype(np.float64), 0, upper).astype(np.int64)
        d = coords - np.floor(coords)
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    with np.errstate(invalid='ignore'):
        Q1 = np.stack([np.ones_like(dx), dx, dy, dz, dx * dy, dy * dz, dx * dz, dx * dy * dz], 0)
        W = (Q1.T.astype(np.float32) @ _B1).astype(np.float32)       # [n, 8]
        P = volume[idx[..., 0], idx[..., 1], idx[..., 2]]            # [n, 8, C]
        out = np.sum(P * W[:, :, None], axis=1, dtype=np.float32)
    return out.astype(np.float32)


def interpolate_in_neighborhood(volume, coords, nb_dirs):
    """dwi_ml ``interpolate_volume_in_neighborhood``: point-major / neighbour-minor flat
    coordinates, result reshaped [n, 7*C] (call site environments/env.py:538-541)."""
    n = coords.shape[0]
    flat = (coords[:, None, :].astype(np.float32) + nb_dirs[None]).reshape(-1, 3)
    return trilinear(volume, flat).reshape(n, -1)


def format_state(volume, streamlines, nb_dirs, n_dirs):
    """environments/env.py:504-565 (_format_state)."""
    N, L, P = streamlines.shape
    tip = streamlines[:, -1, :]
    signal = interpolate_in_neighborhood(volume, tip, nb_dirs)
    S = signal.shape[1]
    inputs = np.zeros((N, S + n_dirs * P), dtype=np.float32)
    inputs[:, :S] = signal
    previous_dirs = np.zeros((N, n_dirs, P), dtype=np.float32)
    if L > 1:
        with np.errstate(invalid='ignore'):
            dirs = np.diff(streamlines, axis=1)
        previous_dirs[:, :min(dirs.shape[1], n_dirs), :] = dirs[:, :-(n_dirs + 1):-1, :]
    inputs[:, S:] = previous_dirs.reshape(N, n_dirs * P)
    return inputs


# --------------------------------------------------------------------------------------
# S6-S8: stopping criteria
# --------------------------------------------------------------------------------------
def is_too_long(streamlines, max_nb_steps):
    """environments/utils.py:127-142."""
    return np.full(streamlines.shape[0], streamlines.shape[1] >= max_nb_steps)


def is_too_curvy(streamlines, max_theta, compare_f32=True):
    """environments/utils.py:145-173.  ``compare_f32`` reproduces numpy 1.23 (the
    reference's pinned version), where ``float32_array > np.float64_scalar`` compares in
    float32; numpy 2 compares in float64 -- they differ only when the angle equals
    float32(theta) exactly."""
    max_theta_rad = np.deg2rad(max_theta)
    if streamlines.shape[1] < 3:
        return np.zeros(streamlines.shape[0], dtype=bool)
    with np.errstate(all='ignore'):
        u = normalize_vectors(streamlines[:, -1] - streamlines[:, -2])
        v = normalize_vectors(streamlines[:, -2] - streamlines[:, -3])
        angles = np.arccos(np.einsum('ij,ij->i', u, v))
        if compare_f32 and angles.dtype == np.float32:
            return angles > np.float32(max_theta_rad)
        return angles > max_theta_rad


class BinaryStoppingCriterion:
    """environments/stopping_criteria.py:38-82 -- real scipy, as in the reference."""

    def __init__(self, mask, threshold=0.5):
        self.mask = spline_filter(np.ascontiguousarray(mask, dtype=float), order=3)
        self.threshold = threshold

    def values(self, streamlines):
        coords = streamlines[:, -1, :].T - 0.5
        return map_coordinates(self.mask, coords, prefilter=False)

    def __call__(self, streamlines):
        return self.values(streamlines) < self.threshold


def bspline3_weights(y):
    """scipy ni_splines.c get_spline_interpolation_weights, order 3 (y = x - floor(x))."""
    z = 1.0 - y
    w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
    w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
    w0 = z * z * z / 6.0
    w3 = 1.0 - w0 - w1 - w2
    return w0, w1, w2, w3


def _mirror(idx, n):
    """scipy NI_GeometricTransform edge handling for mode 'mirror' taps."""
    if n <= 1:
        return 0
    s2 = 2 * n - 2
    if idx < 0:
        idx = s2 * int(-idx / s2) + idx
        return idx + s2 if idx <= 1 - n else -idx
    if idx >= n:
        idx -= s2 * int(idx / s2)
        if idx >= n:
            idx = s2 - idx
    return idx


def spline_mask_value_restated(coef, point_f32):
    """What the CUDA kernel implements for S8, written out tap by tap (pure Python, small
    cases only): scipy ``map_coordinates(coef, p - 0.5, order=3, mode='constant', cval=0,
    prefilter=False)`` = 0 outside [0, n-1] or for NaN/inf, else 64 mirror-indexed taps.
    tests/test_oracle_golden.py checks it against the real scipy call."""
    c = (np.asarray(point_f32, dtype=np.float32) - np.float32(0.5)).astype(np.float64)
    dims = coef.shape
    starts, ws = [], []
    for a in range(3):
        cc = c[a]
        if not (cc >= 0.0 and cc <= dims[a] - 1):
            return 0.0
        fl = math.floor(cc)
        starts.append(int(fl) - 1)
        ws.append(bspline3_weights(cc - fl))
    out = 0.0
    for i in range(4):
        xi = _mirror(starts[0] + i, dims[0])
        for j in range(4):
            yj = _mirror(starts[1] + j, dims[1])
            for k in range(4):
                zk = _mirror(starts[2] + k, dims[2])
                t = coef[xi, yj, zk]
                t *= ws[0][i]
                t *= ws[1][j]
                t *= ws[2][k]
                out += t
    return out


# --------------------------------------------------------------------------------------
# S11-S12: reward
# --------------------------------------------------------------------------------------
def nearest_neighbor_interpolation(volume, coords):
    """environments/interpolation.py:7-26."""
    indices_unclipped = np.round(coords).astype(np.int32)
    upper = (np.asarray(volume.shape[:3]) - 1)
    indices = np.clip(indices_unclipped, 0, upper).astype(int).T
    return volume[tuple(indices)]


def peaks_alignment_reward(peaks, streamlines):
    """environments/local_reward.py:29-107 (PeaksAlignmentReward.__call__)."""
    N, L, _ = streamlines.shape
    if L < 2:
        return np.ones(N, dtype=np.uint8)
    P = peaks.shape[-1]
    with np.errstate(all='ignore'):
        idx = streamlines[:, -2].astype(np.int32)
        v = nearest_neighbor_interpolation(peaks, idx)
        v = np.reshape(v, (N * 5, P // 5))
        v = normalize_vectors(v)
        v = np.reshape(v, (N, 5, P // 5))
        v = np.nan_to_num(v)
        dirs = np.diff(streamlines, axis=1)
        u = np.nan_to_num(normalize_vectors(dirs[:, -1]))
        dot = np.abs(np.einsum('ijk,ik->ij', v, u))
        rewards = np.amax(dot, axis=-1)
        factors = np.ones((N))
        if L >= 3:
            w = np.nan_to_num(normalize_vectors(dirs[:, -2]))
            np.einsum('ik,ik->i', u, w, out=factors)
        rewards *= factors
    return rewards


# --------------------------------------------------------------------------------------
# S1-S5, S10, S15-S16: the environment
# --------------------------------------------------------------------------------------
class OracleEnv:
    """environments/tracking_env.py:13-294 + environments/env.py:143-282,493-603 restated.

    ``noisy=True`` models NoisyTrackingEnvironment (noisy_tracking_env.py:38-77): actions get
    a float64 ``rng.normal(0, noise)`` added, so directions are float64 even at noise 0
    (SURVEY.md F7).  ``noisy=False`` models TrackingEnvironment under the reference's pinned
    numpy 1.23: float32 throughout.
    """

    def __init__(self, sh, mask, seeds, voxel_size, step_size_mm, theta=30., n_dirs=100,
                 max_length_mm=200., min_length_mm=20., threshold=0.1, peaks=None,
                 compute_reward=False, alignment_weighting=1.0, noisy=False, noise=0.0,
                 rng=None, oracle_ckpt=None, oracle_stopping=False, oracle_bonus=0.0):
        self.volume = np.ascontiguousarray(sh, dtype=np.float32)
        self.seeds = np.asarray(seeds)
        self.theta = theta
        self.n_dirs = n_dirs
        self.noisy = noisy
        self.noise = noise
        self.rng = rng or np.random.RandomState(1337)
        # env.py:196-213 / datasets/utils.py:88-124
        self.step_size_mm = step_size_mm
        self.step_size = step_size_mm / np.mean(np.abs([voxel_size] * 3))     # np.float64
        self.max_nb_steps = int(max_length_mm / step_size_mm)
        self.min_nb_steps = int(min_length_mm / step_size_mm)
        self.nb_dirs = neighborhood_directions(self.step_size)
        self.mask_criterion = BinaryStoppingCriterion(np.asarray(mask).astype(np.uint8), threshold)
        self.peaks = peaks
        self.compute_reward = compute_reward
        self.alignment_weighting = alignment_weighting
        self.oracle_ckpt = oracle_ckpt
        self.oracle_stopping = oracle_stopping
        self.oracle_bonus = oracle_bonus

    # ---- env.py:567-603 in the dict order of env.py:233-260 (no oracle criterion here)
    def _is_stopping(self, streamlines):
        n = len(streamlines)
        should_stop = np.zeros(n, dtype=bool)
        flags = np.zeros(n, dtype=int)
        criteria = [(LENGTH, is_too_long(streamlines, self.max_nb_steps)),
                    (CURVATURE, is_too_curvy(streamlines, self.theta))]
        if self.oracle_ckpt is not None and self.oracle_stopping:
            # stopping_criteria.py:113-154 with min_nb_steps * 5 (env.py:246-253)
            if streamlines.shape[1] > self.min_nb_steps * 5:
                criteria.append((ORACLE, oracle_predict(self.oracle_ckpt, list(streamlines)) < 0.5))
        criteria.append((MASK, self.mask_criterion(streamlines)))
        for bit, hit in criteria:
            flags[hit] |= bit
            should_stop[hit] = True
        return should_stop, flags

    def _format_actions(self, actions):
        """env.py:493-502."""
        with np.errstate(all='ignore'):
            if self.noisy:
                return normalize_vectors(actions) * self.step_size              # float64
            a = np.asarray(actions, dtype=np.float32)
            return (normalize_vectors(a) * np.float32(self.step_size)).astype(np.float32)

    def _format_state(self, streamlines):
        return format_state(self.volume, streamlines, self.nb_dirs, self.n_dirs)

    def _start(self, initial_points):
        N = initial_points.shape[0]
        self.initial_points = initial_points
        self.streamlines = np.zeros((N, self.max_nb_steps + 1, 3), dtype=np.float32)
        self.streamlines[:, 0, :] = initial_points
        self.flags = np.zeros(N, dtype=int)
        self.lengths = np.ones(N, dtype=np.int32)
        self.length = 1
        self.dones = np.full(N, False)
        self.continue_idx = np.arange(N)
        self.state = self._format_state(self.streamlines[self.continue_idx, :self.length])
        return self.state[self.continue_idx]

    def reset(self, start, end):
        """tracking_env.py:91-133."""
        return self._start(self.seeds[start:end])

    def nreset(self, n_seeds):
        """tracking_env.py:47-89."""
        replace = n_seeds > len(self.seeds)
        sel = np.random.choice(np.arange(len(self.seeds)), size=n_seeds, replace=replace)
        return self._start(self.seeds[sel])

    def step(self, actions):
        """tracking_env.py:135-221 (+ noisy_tracking_env.py:63-77)."""
        if self.noisy:
            actions = actions + self.rng.normal(0., self.noise, size=actions.shape)
        directions = self._format_actions(actions)
        ci = self.continue_idx
        with np.errstate(all='ignore'):
            if self.length == 1:
                streamlines = np.array(self.streamlines[ci])
                streamlines[:, self.length, :] = self.streamlines[ci, self.length - 1, :] + directions
                stopping, _ = self._is_stopping(streamlines[:, :self.length + 1])
                directions[stopping] *= -1
            self.streamlines[ci, self.length, :] = self.streamlines[ci, self.length - 1, :] + directions
        self.length += 1
        stopping, new_flags = self._is_stopping(self.streamlines[ci, :self.length])
        self.not_stopping = np.logical_not(stopping)
        self.new_continue_idx, self.stopping_idx = ci[~stopping], ci[stopping]
        self.flags[self.stopping_idx] = new_flags[stopping]
        self.dones[self.stopping_idx] = 1
        reward = np.zeros(self.streamlines.shape[0])
        if self.compute_reward:
            # reward.py:46-79 with factors [peaks (w), oracle (0)]
            reward = self.alignment_weighting * peaks_alignment_reward(
                self.peaks, self.streamlines[ci, :self.length]).astype(np.float64)
            if self.oracle_ckpt is not None and self.oracle_bonus > 0:
                # oracle_reward.py:70-93: sparse bonus for rows done this step that the oracle likes
                dn = self.dones[ci]
                if self.length > self.min_nb_steps and dn.sum() > 0:
                    pred = oracle_predict(self.oracle_ckpt, list(self.streamlines[ci, :self.length][dn]))
                    bonus = np.zeros(len(ci))
                    bonus[np.arange(len(ci))[dn][pred > 0.5]] = 1.0
                    reward = reward + self.oracle_bonus * bonus
        self.state[ci] = self._format_state(self.streamlines[ci, :self.length])
        return self.state[ci], reward, self.dones[ci], {'continue_idx': ci}

    def harvest(self):
        """tracking_env.py:223-245."""
        self.lengths[self.stopping_idx] = self.length
        self.continue_idx = self.new_continue_idx
        return self.state[self.continue_idx], self.not_stopping

    def get_streamlines(self):
        """tracking_env.py:247-294: list of [len,3] float32, seeds, flags."""
        out = [self.streamlines[i, :self.lengths[i], :] for i in range(len(self.streamlines))]
        cut = np.logical_or(is_flag_set(self.flags, CURVATURE), is_flag_set(self.flags, MASK))
        out = [s[:-1] if f else s for s, f in zip(out, cut)]
        return out, self.initial_points, self.flags


# --------------------------------------------------------------------------------------
# A1: SAC actor forward
# --------------------------------------------------------------------------------------
def actor_forward(sd, state, probabilistic=0.0, eps=None):
    """algorithms/shared/offpolicy.py:94-140 (MaxEntropyActor.forward) with the network of
    algorithms/shared/utils.py:41-51, float32 numpy.  ``sd``: layers.{0,2,..}.{weight,bias}.
    ``eps`` is the N(0,1) draw of ``Normal.rsample``.  Returns (action, logp, pre-activation)."""
    h = np.asarray(state, dtype=np.float32)
    n_layers = len([k for k in sd if k.endswith('.weight')])
    for li in range(n_layers):
        w = np.asarray(sd['layers.%d.weight' % (2 * li)], dtype=np.float32)
        b = np.asarray(sd['layers.%d.bias' % (2 * li)], dtype=np.float32)
        h = h @ w.T + b
        if li < n_layers - 1:
            h = np.maximum(h, 0)
    p = h
    A = p.shape[1] // 2
    mu = p[:, :A]
    log_std = np.clip(p[:, A:], -20, 2)
    std = np.exp(log_std) * np.float32(probabilistic)
    if eps is None:
        eps = np.zeros_like(mu)
    pi = mu + std * eps
    with np.errstate(all='ignore'):
        # Normal.log_prob: -((x-mu)^2)/(2 var) - log(std) - log(sqrt(2 pi))
        logp = (-((pi - mu) ** 2) / (2 * std * std) - np.log(std)
                - np.float32(math.log(math.sqrt(2 * math.pi)))).sum(-1)
        softplus = np.logaddexp(0, -2 * pi)
        logp = logp - (2 * (np.log(2) - pi - softplus)).sum(1)
    return np.tanh(pi).astype(np.float32), logp.astype(np.float32), p


# --------------------------------------------------------------------------------------
# L1: load-time peak extraction (environments/env.py:405-432)
# --------------------------------------------------------------------------------------
def sh_basis_matrix(vertices, order=8):
    """dipy ``sh_to_sf_matrix(sphere, order, "descoteaux07")`` (legacy basis, env.py:414) restated with
    scipy's complex harmonics: for even l and m = -l..l, sqrt(2) Re Y_l^|m| (m < 0), Y_l^0,
    sqrt(2) Im Y_l^m (m > 0).  -> B [V, n_coefs] float64, SF = sh . B^T."""
    from scipy.special import sph_harm_y
    v = np.asarray(vertices, dtype=np.float64)
    polar = np.arccos(np.clip(v[:, 2], -1.0, 1.0))
    azim = np.arctan2(v[:, 1], v[:, 0])
    cols = []
    for l in range(0, order + 1, 2):
        for m in range(-l, l + 1):
            y = sph_harm_y(l, abs(m), polar, azim)
            if m < 0:
                cols.append(np.sqrt(2.0) * y.real)
            elif m == 0:
                cols.append(y.real)
            else:
                cols.append(np.sqrt(2.0) * y.imag)
    return np.stack(cols, axis=1)


def local_maxima(odf, edges):
    """dipy ``local_maxima`` (reconst/recspeed.pyx, restated): a vertex is a peak when it is greater
    than at least one neighbour and smaller than none.  Values descending (ties: lower index first)."""
    odf = np.asarray(odf, dtype=np.float64)
    state = np.zeros(len(odf), dtype=np.int64)        # 0 unvisited, 1 maybe, 2 not a peak
    for a, b in edges:
        if odf[a] < odf[b]:
            state[a] = 2
            state[b] = max(state[b], 1)
        elif odf[a] > odf[b]:
            state[a] = max(state[a], 1)
            state[b] = 2
    idx = np.nonzero(state == 1)[0]
    order = np.lexsort((idx, -odf[idx]))
    idx = idx[order]
    return odf[idx], idx


def peak_directions(odf, vertices, edges, relative_peak_threshold=0.5, min_separation_angle=25.0):
    """dipy ``peak_directions`` restated: local maxima, drop those below the relative threshold
    (measured from max(min(odf), 0)), then greedily drop directions within the separation angle
    (|cos|, antipodal symmetry) of an already kept one."""
    values, indices = local_maxima(odf, edges)
    n = len(values)
    if n == 0 or values[0] < 0.0:
        return np.zeros((0, 3)), np.zeros(0), np.zeros(0, dtype=np.int64)
    if n == 1:
        return vertices[indices], values, indices
    odf_min = max(float(np.min(odf)), 0.0)
    norm = values - odf_min
    n = int(np.sum(np.cumprod(norm >= relative_peak_threshold * norm[0])))
    indices = indices[:n]
    dirs = vertices[indices]
    cos_sim = np.cos(np.deg2rad(min_separation_angle))
    kept = []
    for i in range(n):
        if all(abs(float(np.dot(dirs[i], dirs[j]))) <= cos_sim for j in kept):
            kept.append(i)
    kept = np.asarray(kept, dtype=np.int64)
    return dirs[kept], values[kept], indices[kept]


def peaks_from_sh(data, vertices, edges, B, npeaks=5, relative_threshold=0.1, absolute_threshold=0.0):
    """environments/env.py:405-432: per voxel with a non-zero coefficient sum, scilpy
    ``get_maximas(sh, sphere, B, 0.1, 0)`` (SF = sh . B^T, values below the absolute threshold zeroed,
    ``peak_directions`` with the default 25 degree separation), the first ``npeaks`` directions scaled by
    value / first value -> [..., npeaks * 3] float32 (``reshape_peaks_for_visualization``)."""
    data = np.asarray(data)
    shape = data.shape[:-1]
    flat = data.reshape(-1, data.shape[-1])
    out = np.zeros((flat.shape[0], npeaks, 3), dtype=np.float64)
    for i in np.nonzero(np.sum(flat, axis=-1))[0]:
        sf = np.dot(flat[i], B.T)
        sf[sf < absolute_threshold] = 0.0
        d, val, _ = peak_directions(sf, vertices, edges, relative_threshold, 25.0)
        if len(val):
            n = min(npeaks, len(val))
            w = val[:n] / val[0] if val[0] != 0 else np.zeros(n)
            out[i, :n] = d[:n] * w[:, None]
    return out.reshape(shape + (npeaks * 3,)).astype(np.float32)


# --------------------------------------------------------------------------------------
# O1-O2: TractOracle-Net
# --------------------------------------------------------------------------------------
def streamline_length(s):
    """dipy ``length`` (tracking/tracker.py:120): sum of segment norms."""
    s = np.asarray(s, dtype=np.float64)
    if len(s) < 2:
        return 0.0
    return float(np.sqrt(((s[1:] - s[:-1]) ** 2).sum(-1)).sum())


def compress_streamline(s, tol_error=0.01, max_segment_length=10.0):
    """dipy ``compress_streamlines`` for one streamline (tracking/tracker.py:123-125, ``--compress``),
    restated from the published algorithm of dipy/tracking/streamlinespeed.pyx
    (``c_compress_streamline`` / ``c_dist_to_line`` / ``c_segment_length``; dipy is not installed here
    and the reference does not pin its version, so this restatement is the definition the CUDA
    kernel is held to -- parity with upstream unpinned):

      * the first and last points are kept; streamlines of <= 2 points are copied;
      * walking ``nxt`` = 2..N-1 with ``prev`` the last kept point: the chord (prev, nxt) replaces
        the path if it is shorter than ``max_segment_length`` and every point strictly between them
        lies within ``tol_error`` of the line through prev and nxt (a NaN distance counts as too
        far); otherwise point nxt-1 is kept and becomes ``prev``;
      * distances: |(nxt-prev) x (curr-nxt)| / |nxt-prev| with coordinate differences and their
        products in the streamline's dtype (C float arithmetic) summed in double;
        chord length: double sum of squared (float) differences.
    Returns the compressed streamline (same dtype)."""
    s = np.asarray(s)
    N = len(s)
    if N <= 2:
        return s.copy()
    dt = s.dtype.type
    keep = [0]
    prev = 0
    for nxt in range(2, N):
        dn = (s[nxt] - s[prev]).astype(np.float64)              # float difference widened to double
        seg = np.sqrt(dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2])
        ok = False
        if seg < max_segment_length:
            ok = True
            a = s[nxt] - s[prev]                                 # dtype arithmetic
            norm2 = np.sqrt(np.float64(dt(a[0] * a[0])) + np.float64(dt(a[1] * a[1])) + np.float64(dt(a[2] * a[2])))
            for curr in range(prev + 1, nxt):
                b = s[curr] - s[nxt]
                cx = np.float64(dt(dt(a[1] * b[2]) - dt(a[2] * b[1])))
                cy = np.float64(dt(dt(a[2] * b[0]) - dt(a[0] * b[2])))
                cz = np.float64(dt(dt(a[0] * b[1]) - dt(a[1] * b[0])))
                with np.errstate(invalid='ignore', divide='ignore'):
                    dist = np.sqrt(cx * cx + cy * cy + cz * cz) / norm2
                if np.isnan(dist) or dist > tol_error:
                    ok = False
                    break
        if not ok:
            keep.append(nxt - 1)
            prev = nxt - 1
    keep.append(N - 1)
    return s[np.asarray(keep)]


def set_number_of_points(s, nb_points=128):
    """dipy ``set_number_of_points`` restated (oracles/oracle.py:52,70):
    arc-length linear resampling; differences in the input dtype, arc lengths and
    interpolation in double, stored back in the input dtype, last point copied."""
    s = np.asarray(s)
    N = len(s)
    d = np.diff(s, axis=0).astype(np.float64)
    seg = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    cum = np.zeros(N, dtype=np.float64)
    for i in range(1, N):
        cum[i] = cum[i - 1] + seg[i - 1]
    step = cum[N - 1] / (nb_points - 1)
    res = np.zeros((nb_points, 3), dtype=s.dtype)
    nxt, i, k = 0.0, 0, 0
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
    return res


def oracle_features(streamlines, nb_points=128):
    """oracles/oracle.py:52-54: resample to 128 points, np.diff -> [B,127,3] float32."""
    data = np.stack([set_number_of_points(np.asarray(s, dtype=np.float32), nb_points)
                     for s in streamlines])
    return np.diff(data, axis=1).astype(np.float32)


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def transformer_oracle_forward(ckpt, dirs):
    """oracles/transformer_oracle.py:77-92 restated in float32 numpy: CLS token prepend,
    relu(Linear(3,32))*sqrt(32), + sinusoidal PE, n_layers post-norm encoder layers
    (torch nn.TransformerEncoderLayer defaults: relu, dim_feedforward 2048, eps 1e-5,
    eval mode so dropout is off), sigmoid(Linear(32,1)) of token 0."""
    hp = ckpt['hyper_parameters']
    sd = {k: np.asarray(v, dtype=np.float32) for k, v in ckpt['state_dict'].items()}
    n_head, n_layers = hp['n_head'], hp['n_layers']
    x = np.asarray(dirs, dtype=np.float32)
    B = x.shape[0]
    cls = np.broadcast_to(sd['cls_token'][None, None, :], (B, 1, 3))
    x = np.concatenate((cls, x), axis=1)                                   # [B,128,3]
    E = sd['embedding.0.weight'].shape[0]
    x = np.maximum(x @ sd['embedding.0.weight'].T + sd['embedding.0.bias'], 0) * np.float32(math.sqrt(E))
    T = x.shape[1]
    x = x + sd['pos_encoding.pe'][:T, 0][None]
    dh = E // n_head
    for i in range(n_layers):
        p = 'bert.layers.%d.' % i
        qkv = x @ sd[p + 'self_attn.in_proj_weight'].T + sd[p + 'self_attn.in_proj_bias']
        q, k, v = np.split(qkv, 3, axis=-1)

        def heads(t):
            return t.reshape(B, T, n_head, dh).transpose(0, 2, 1, 3)
        q, k, v = heads(q), heads(k), heads(v)
        s = (q @ k.transpose(0, 1, 3, 2)) / np.float32(math.sqrt(dh))
        s = s - s.max(-1, keepdims=True)
        a = np.exp(s)
        a = a / a.sum(-1, keepdims=True)
        o = (a @ v).transpose(0, 2, 1, 3).reshape(B, T, E)
        o = o @ sd[p + 'self_attn.out_proj.weight'].T + sd[p + 'self_attn.out_proj.bias']
        x = _layer_norm(x + o, sd[p + 'norm1.weight'], sd[p + 'norm1.bias'])
        f = np.maximum(x @ sd[p + 'linear1.weight'].T + sd[p + 'linear1.bias'], 0)
        f = f @ sd[p + 'linear2.weight'].T + sd[p + 'linear2.bias']
        x = _layer_norm(x + f, sd[p + 'norm2.weight'], sd[p + 'norm2.bias'])
    y = x[:, 0] @ sd['head.weight'].T + sd['head.bias']
    return (1.0 / (1.0 + np.exp(-y)))[:, 0].astype(np.float32)


def oracle_predict(ckpt, streamlines, batch_size=4096):
    """OracleSingleton.predict semantics as used correctly by
    experiment/oracle_validator.py:40-47 (chunks <= 4096; the reference's own loop drops a
    trailing partial batch when N > 4096, SURVEY.md F13 -- not replicated)."""
    out = np.zeros(len(streamlines), dtype=np.float32)
    for i in range(0, len(streamlines), batch_size):
        out[i:i + batch_size] = transformer_oracle_forward(
            ckpt, oracle_features(streamlines[i:i + batch_size]))
    return out


# --------------------------------------------------------------------------------------
# A3 / T1: the episode loop, used as the timed CPU baseline
# --------------------------------------------------------------------------------------
def validation_episode(env, actor_sd, start, end, torch_actor=None):
    """algorithms/rl.py:58-106 with prob = 0 (tracking/tracker.py:28).  Returns the number
    of streamline-steps taken.  ``torch_actor`` (a callable state->action on torch CPU,
    all host threads) replaces the numpy actor when given -- it is what the reference runs."""
    state = env.reset(start, end)
    steps = 0
    done = np.array([False])
    while not np.all(done):
        if torch_actor is not None:
            action = torch_actor(state)
        else:
            action, _, _ = actor_forward(actor_sd, state, 0.0)
        steps += len(env.continue_idx)
        _, _, done, _ = env.step(action)
        state, _ = env.harvest()
    return steps
