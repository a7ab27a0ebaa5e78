"""Seeded synthetic inputs for the tracking hot path.

The reference ships no fixture data, so every test, the smoke
run and bench.py build their inputs here: an order-8 descoteaux07 fODF volume made
from an analytic fibre field, an ellipsoid tracking mask, a 1-voxel shell seeding
mask, matching peaks for the alignment reward, seeds drawn the way dipy's
``random_seeds_from_mask`` draws them (reference call: environments/env.py:216-219),
and actor / TractOracle-Net checkpoints written in the reference's formats
(algorithms/shared/offpolicy.py:327-340, oracles/transformer_oracle.py:95-118).

Everything is plain torch/numpy on whatever device is asked for; nothing here is on
the timed path.
"""
import json
import math
import os

import numpy as np
import torch

SH_ORDER = 8
N_COEFS = 45


# ---------------------------------------------------------------------------
# Real, symmetric spherical harmonics (descoteaux07 "legacy" convention):
#   coefficient index runs l = 0,2,..,8 ; m = -l..l
#   m < 0 : sqrt(2) * Re(Y_l^|m|) ; m = 0 : Y_l^0 ; m > 0 : sqrt(2) * Im(Y_l^m)
# ---------------------------------------------------------------------------
def _legendre_table(lmax, x):
    """Associated Legendre P_l^m(x) for 0<=m<=l<=lmax (with Condon-Shortley phase)."""
    P = {}
    somx2 = torch.sqrt(torch.clamp(1.0 - x * x, min=0.0))
    P[(0, 0)] = torch.ones_like(x)
    for m in range(1, lmax + 1):
        P[(m, m)] = -(2 * m - 1) * somx2 * P[(m - 1, m - 1)]
    for m in range(0, lmax):
        P[(m + 1, m)] = (2 * m + 1) * x * P[(m, m)]
    for m in range(0, lmax + 1):
        for l in range(m + 2, lmax + 1):
            P[(l, m)] = ((2 * l - 1) * x * P[(l - 1, m)]
                         - (l + m - 1) * P[(l - 2, m)]) / (l - m)
    return P


def real_sh_basis(dirs, order=SH_ORDER):
    """dirs: [..., 3] unit vectors -> [..., n_coefs] real even SH basis values."""
    dirs = dirs.to(torch.float64)
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    ct = torch.clamp(z, -1.0, 1.0)
    phi = torch.atan2(y, x)
    P = _legendre_table(order, ct)
    out = []
    for l in range(0, order + 1, 2):
        for m in range(-l, l + 1):
            am = abs(m)
            norm = math.sqrt((2 * l + 1) / (4 * math.pi)
                             * math.factorial(l - am) / math.factorial(l + am))
            if m < 0:
                out.append(math.sqrt(2.0) * norm * P[(l, am)] * torch.cos(am * phi))
            elif m == 0:
                out.append(norm * P[(l, 0)])
            else:
                out.append(math.sqrt(2.0) * norm * P[(l, am)] * torch.sin(am * phi))
    return torch.stack(out, dim=-1)


def _zonal_response(order=SH_ORDER):
    r = []
    for l in range(0, order + 1, 2):
        r.extend([math.exp(-l * (l + 1) / 12.0)] * (2 * l + 1))
    return torch.tensor(r, dtype=torch.float64)


def fibre_field(shape, device='cpu'):
    """Analytic two-population fibre field.

    Returns (u1 [X,Y,Z,3], u2 [X,Y,Z,3], f2 [X,Y,Z]) with f2 > 0 only in the
    central third (a 60-degree crossing), all float64.
    """
    X, Y, Z = shape
    gx = torch.arange(X, dtype=torch.float64, device=device)[:, None, None]
    gy = torch.arange(Y, dtype=torch.float64, device=device)[None, :, None]
    gz = torch.arange(Z, dtype=torch.float64, device=device)[None, None, :]
    phi = 2 * math.pi * (gy / Y + 0.25 * torch.sin(2 * math.pi * gz / Z)) + 0 * gx
    u1 = torch.stack((torch.cos(phi), torch.sin(phi), torch.full_like(phi, 0.3)), -1)
    u1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
    phi2 = phi + math.pi / 3
    u2 = torch.stack((torch.cos(phi2), torch.sin(phi2), torch.full_like(phi, 0.3)), -1)
    u2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
    central = ((gx > X / 3) & (gx < 2 * X / 3) & (gy > Y / 3) & (gy < 2 * Y / 3)
               & (gz > Z / 3) & (gz < 2 * Z / 3))
    f2 = torch.where(central, 0.6, 0.0).to(torch.float64)
    return u1, u2, f2


def ellipsoid_mask(shape, frac=0.42, device='cpu'):
    X, Y, Z = shape
    gx = (torch.arange(X, dtype=torch.float64, device=device)[:, None, None] - (X - 1) / 2) / (frac * X)
    gy = (torch.arange(Y, dtype=torch.float64, device=device)[None, :, None] - (Y - 1) / 2) / (frac * Y)
    gz = (torch.arange(Z, dtype=torch.float64, device=device)[None, None, :] - (Z - 1) / 2) / (frac * Z)
    return (gx * gx + gy * gy + gz * gz) <= 1.0


def shell_mask(mask):
    """1-voxel-thick shell at the inner surface of a boolean mask (6-connectivity)."""
    m = mask
    inner = m.clone()
    for d in range(3):
        lo = torch.roll(m, 1, d)
        hi = torch.roll(m, -1, d)
        idx_lo = [slice(None)] * 3
        idx_hi = [slice(None)] * 3
        idx_lo[d] = 0
        idx_hi[d] = -1
        lo[tuple(idx_lo)] = False
        hi[tuple(idx_hi)] = False
        inner &= lo & hi
    return m & ~inner


def make_subject(shape, seed=1234, noise=0.01, device='cpu', with_peaks=True,
                 chunk=32):
    """Synthetic subject: dict with

    sh     [X,Y,Z,45] float32  descoteaux07 fODF coefficients (zero outside the mask)
    mask   [X,Y,Z]    uint8    tracking mask (ellipsoid)
    seed_mask [X,Y,Z] uint8    seeding mask (ellipsoid shell)
    peaks  [X,Y,Z,15] float32  up to 5 value-normalised peak directions (2 used)
    """
    X, Y, Z = shape
    mask = ellipsoid_mask(shape, device=device)
    seed_mask = shell_mask(mask)
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    resp = _zonal_response().to(device)
    sh = torch.empty((X, Y, Z, N_COEFS), dtype=torch.float32, device=device)
    peaks = torch.zeros((X, Y, Z, 15), dtype=torch.float32, device=device) if with_peaks else None
    u1, u2, f2 = fibre_field(shape, device=device)
    for x0 in range(0, X, chunk):
        sl = slice(x0, min(X, x0 + chunk))
        b1 = real_sh_basis(u1[sl]) * resp
        b2 = real_sh_basis(u2[sl]) * resp
        c = b1 + f2[sl][..., None] * b2
        nz = torch.randn(c.shape, generator=g, dtype=torch.float32).to(device)
        c = c.to(torch.float32) + noise * nz
        c[..., 0] = torch.abs(c[..., 0])
        c = c * mask[sl][..., None]
        sh[sl] = c
        if with_peaks:
            m = mask[sl][..., None].to(torch.float32)
            peaks[sl, :, :, 0:3] = u1[sl].to(torch.float32) * m
            peaks[sl, :, :, 3:6] = (u2[sl] * f2[sl][..., None]).to(torch.float32) * m
    return {
        'sh': sh,
        'mask': mask.to(torch.uint8),
        'seed_mask': seed_mask.to(torch.uint8),
        'peaks': peaks,
    }


def seeds_from_mask(seed_mask, npv, rng):
    """dipy ``random_seeds_from_mask(mask, eye(4), seeds_count=npv)`` restated.

    dipy (unpinned transitive dependency of the reference, call site
    environments/env.py:216-219) loops ``for i in 1..npv: for s in argwhere(mask):
    s + np.random.random(3) - 0.5``; ``rng.random_sample((npv, n, 3))`` consumes the
    same stream in the same order.  ``rng`` is a ``np.random.RandomState`` (the
    reference uses the global numpy state).  Returns float64 [npv*n, 3] in voxel
    space, voxel centres at integer coordinates.
    """
    where = np.argwhere(np.asarray(seed_mask, dtype=bool))
    grid = rng.random_sample((npv, len(where), 3))
    seeds = where[None, :, :].astype(np.float64) + grid - 0.5
    return seeds.reshape(-1, 3)


# ---------------------------------------------------------------------------
# Checkpoints in the reference's formats
# ---------------------------------------------------------------------------
def actor_state_dict(input_size=615, hidden_dims='1024-1024-1024', action_size=3,
                     seed=1111, kind='random'):
    """SAC MaxEntropyActor state dict: layers.{0,2,4,6}.{weight,bias}
    (algorithms/shared/utils.py:41-51, offpolicy.py:91-92).

    kind='random'  : torch default nn.Linear init under manual_seed(seed) -- parity.
    kind='tracking': random init rescaled + a carry path so that
                     mu ~= previous direction + small SH-driven term; streamlines then
                     survive the 30-degree curvature criterion for hundreds of steps
                     (throughput runs).
    """
    widths = [int(w) for w in hidden_dims.split('-')]
    dims = [input_size] + widths + [2 * action_size]
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for li in range(len(dims) - 1):
        fan_in, fan_out = dims[li], dims[li + 1]
        bound = 1.0 / math.sqrt(fan_in)
        w = (torch.rand((fan_out, fan_in), generator=g) * 2 - 1) * bound
        b = (torch.rand((fan_out,), generator=g) * 2 - 1) * bound
        sd['layers.%d.weight' % (2 * li)] = w
        sd['layers.%d.bias' % (2 * li)] = b
    if kind == 'tracking':
        n_sig = input_size - 300 if input_size > 300 else 0
        # Shrink the random part, then wire relu(+x) / relu(-x) carry units through every
        # hidden layer for the most recent direction (state[n_sig:n_sig+3]) and for a fixed
        # readout of three centre-point SH coefficients (gives the first step a direction and
        # later steps a gentle curvature).  The output gain keeps |mu| <= ~0.35 so tanh stays
        # near-linear and the direction is carried almost unchanged.
        for li in range(len(dims) - 1):
            sd['layers.%d.weight' % (2 * li)] *= 0.02
            sd['layers.%d.bias' % (2 * li)] *= 0.02
        w0 = sd['layers.0.weight']
        for a in range(3):
            for s, sign in enumerate((1.0, -1.0)):
                u = 2 * a + s
                w0[u, :] = 0
                sd['layers.0.bias'][u] = 0
                if n_sig:
                    w0[u, n_sig + a] = sign
                    w0[u, 1 + a] = sign * 0.125
        for li in range(1, len(dims) - 2):
            w = sd['layers.%d.weight' % (2 * li)]
            for u in range(6):
                w[u, :] = 0
                w[:, u] = 0
                w[u, u] = 1.0
                sd['layers.%d.bias' % (2 * li)][u] = 0
        wl = sd['layers.%d.weight' % (2 * (len(dims) - 2))]
        for a in range(3):
            wl[a, :6] = 0
            wl[a, 2 * a] = 0.4375
            wl[a, 2 * a + 1] = -0.4375
    return sd


def critic_state_dict(input_size=615, hidden_dims='1024-1024-1024', action_size=3, seed=1112):
    """DoubleCritic state dict q1/q2.{0,2,4,6}.{weight,bias} (offpolicy.py:213-216)."""
    widths = [int(w) for w in hidden_dims.split('-')]
    dims = [input_size + action_size] + widths + [1]
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for q in ('q1', 'q2'):
        for li in range(len(dims) - 1):
            bound = 1.0 / math.sqrt(dims[li])
            sd['%s.%d.weight' % (q, 2 * li)] = (torch.rand((dims[li + 1], dims[li]), generator=g) * 2 - 1) * bound
            sd['%s.%d.bias' % (q, 2 * li)] = (torch.rand((dims[li + 1],), generator=g) * 2 - 1) * bound
    return sd


DEFAULT_HYPERPARAMETERS = {
    'algorithm': 'SACAuto', 'step_size': 0.75, 'voxel_size': '0.9987237', 'max_angle': 30,
    'hidden_dims': '1024-1024-1024', 'n_dirs': 100, 'target_sh_order': 8.0,
    'input_size': 615, 'action_size': 3, 'min_length': 20.0, 'max_length': 200.0,
    'n_actor': 4096, 'binary_stopping_threshold': 0.1, 'noise': 0.0, 'prob': 1.0,
}


def write_agent_dir(path, kind='random', hidden_dims='1024-1024-1024', input_size=615,
                    seed=1111, hyperparameters=None):
    """Write <path>/last_model_state_{actor,critic}.pth + hyperparameters.json the way
    trainers/train.py:151-179 does."""
    os.makedirs(path, exist_ok=True)
    torch.save(actor_state_dict(input_size, hidden_dims, seed=seed, kind=kind),
               os.path.join(path, 'last_model_state_actor.pth'))
    torch.save(critic_state_dict(input_size, hidden_dims, seed=seed + 1),
               os.path.join(path, 'last_model_state_critic.pth'))
    hp = dict(DEFAULT_HYPERPARAMETERS)
    hp['hidden_dims'] = hidden_dims
    hp['input_size'] = input_size
    if hyperparameters:
        hp.update(hyperparameters)
    with open(os.path.join(path, 'hyperparameters.json'), 'w') as f:
        json.dump(hp, f, indent=4)
    return path


def oracle_checkpoint(n_head=4, n_layers=4, input_size=384, d_model=32, d_ff=2048, seed=2222):
    """TractOracle-Net checkpoint dict {hyper_parameters, state_dict}
    (oracles/oracle.py:21-32, transformer_oracle.py:40-65,95-118)."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd = {'cls_token': torch.randn((3,), generator=g)}
    sd['embedding.0.weight'] = u((d_model, 3), 1 / math.sqrt(3))
    sd['embedding.0.bias'] = u((d_model,), 1 / math.sqrt(3))
    max_len = input_size // 3 + 1
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, 1, d_model)
    pe[:, 0, 0::2] = torch.sin(position * div_term)
    pe[:, 0, 1::2] = torch.cos(position * div_term)
    sd['pos_encoding.pe'] = pe
    for i in range(n_layers):
        p = 'bert.layers.%d.' % i
        sd[p + 'self_attn.in_proj_weight'] = u((3 * d_model, d_model), math.sqrt(6 / (4 * d_model)))
        sd[p + 'self_attn.in_proj_bias'] = torch.zeros(3 * d_model)
        sd[p + 'self_attn.out_proj.weight'] = u((d_model, d_model), 1 / math.sqrt(d_model))
        sd[p + 'self_attn.out_proj.bias'] = torch.zeros(d_model)
        sd[p + 'linear1.weight'] = u((d_ff, d_model), 1 / math.sqrt(d_model))
        sd[p + 'linear1.bias'] = u((d_ff,), 1 / math.sqrt(d_model))
        sd[p + 'linear2.weight'] = u((d_model, d_ff), 1 / math.sqrt(d_ff))
        sd[p + 'linear2.bias'] = u((d_model,), 1 / math.sqrt(d_ff))
        sd[p + 'norm1.weight'] = 1 + 0.1 * u((d_model,), 1.0)
        sd[p + 'norm1.bias'] = 0.1 * u((d_model,), 1.0)
        sd[p + 'norm2.weight'] = 1 + 0.1 * u((d_model,), 1.0)
        sd[p + 'norm2.bias'] = 0.1 * u((d_model,), 1.0)
    sd['head.weight'] = u((1, d_model), 1 / math.sqrt(d_model))
    sd['head.bias'] = u((1,), 1 / math.sqrt(d_model))
    return {
        'hyper_parameters': {'name': 'TransformerOracle', 'input_size': input_size,
                             'output_size': 1, 'lr': 1e-4, 'n_head': n_head,
                             'n_layers': n_layers},
        'state_dict': sd,
    }


def random_streamlines(n, rng, min_pts=20, max_pts=267, step=0.75, start_box=(20., 44.)):
    """Ragged smooth random-walk streamlines (list of [L_i,3] float32) for scoring tests."""
    out = []
    for _ in range(n):
        L = int(rng.randint(min_pts, max_pts + 1))
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        pts = np.empty((L, 3), dtype=np.float32)
        p = rng.uniform(start_box[0], start_box[1], size=3)
        for k in range(L):
            pts[k] = p
            d = d + 0.15 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            p = p + step * d
        out.append(pts)
    return out
