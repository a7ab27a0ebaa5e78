"""tcgen05 dense layer and the SAC actor forward against torch fp32 / the oracle / the
reference-recorded fixture."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O
from tests.helpers import load_golden
from tracktolearn_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (300, 512, 640), (1000, 1024, 1024), (77, 64, 128)])
def test_gemm_bf16_tcgen05(m, n, k):
    from tracktolearn_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(m + n + k)
    A = (torch.randn((m, k), generator=g) * 0.5).cuda().to(torch.bfloat16)
    W = (torch.randn((n, k), generator=g) * 0.1).cuda().to(torch.bfloat16)
    bias = torch.zeros(((n + 255) // 256 * 256,), device='cuda')
    bias[:n] = torch.randn((n,), generator=g).cuda()
    C = torch.zeros((m, n), device='cuda', dtype=torch.bfloat16)
    m_dev = torch.tensor([m], dtype=torch.int32, device='cuda')
    for relu in (0, 1):
        _lib.check(lib.ttl_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(C), m, n, k, n, relu,
                                     _lib.ptr(m_dev), _lib.stream_ptr(torch.device('cuda:0'))), 'gemm')
        torch.cuda.synchronize()
        ref = A.float() @ W.float().t() + bias[:n]
        if relu:
            ref = torch.relu(ref)
        err = (C.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 1e-2 * scale, (err, scale)   # bf16 output rounding (2^-8 relative)


def test_gemm_respects_device_row_count():
    from tracktolearn_b200 import _lib
    lib = _lib.load()
    m, n, k = 512, 256, 128
    A = torch.randn((m, k), device='cuda').to(torch.bfloat16)
    W = torch.randn((n, k), device='cuda').to(torch.bfloat16)
    bias = torch.zeros((256,), device='cuda')
    C = torch.full((m, n), -7.0, device='cuda', dtype=torch.bfloat16)
    m_dev = torch.tensor([130], dtype=torch.int32, device='cuda')
    _lib.check(lib.ttl_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(C), m, n, k, n, 0,
                                 _lib.ptr(m_dev), _lib.stream_ptr(torch.device('cuda:0'))), 'gemm')
    torch.cuda.synchronize()
    assert (C[130:] == -7.0).all()
    ref = A[:130].float() @ W.float().t()
    assert (C[:130].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


def _actor(hidden, seed, precision, kind='random'):
    from tracktolearn_b200.algorithms.shared.offpolicy import SACActorCritic
    sd = synthetic.actor_state_dict(615, hidden, seed=seed, kind=kind)
    ac = SACActorCritic(615, 3, hidden, torch.device('cuda:0'), precision=precision)
    ac.actor.load_state_dict(sd)
    return ac, {k: v.numpy() for k, v in sd.items()}


def test_actor_fp32_tier_matches_reference_fixture():
    g = load_golden('actor')
    hidden = '-'.join(str(int(h)) for h in g['hidden'])
    ac, _ = _actor(hidden, int(g['seed']), 'fp32')
    state = torch.from_numpy(g['state']).cuda()
    a, lp, pre = ac.actor.forward_device(state, 0.0, want_pre=True)
    np.testing.assert_allclose(pre.cpu().numpy(), g['pre'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a.cpu().numpy(), g['action_det'], rtol=1e-5, atol=1e-6)
    a1, lp1, _ = ac.actor.forward_device(state, 1.0, eps=torch.from_numpy(g['eps']).cuda())
    np.testing.assert_allclose(a1.cpu().numpy(), g['action_prob1'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp1.cpu().numpy(), g['logp_prob1'], rtol=1e-4, atol=1e-4)
    assert torch.equal(ac.select_action(state, 0.0), a)


def test_actor_bf16_tier_small_net():
    g = load_golden('actor')
    hidden = '-'.join(str(int(h)) for h in g['hidden'])
    ac, _ = _actor(hidden, int(g['seed']), 'bf16')
    a, _, pre = ac.actor.forward_device(torch.from_numpy(g['state']).cuda(), 0.0, want_pre=True)
    scale = np.abs(g['pre']).max()
    assert np.abs(pre.cpu().numpy() - g['pre']).max() <= 1e-2 * scale
    assert np.abs(a.cpu().numpy() - g['action_det']).max() <= 1e-2


@pytest.mark.parametrize('kind', ['random', 'tracking'])
def test_actor_full_size_both_tiers(kind):
    """615-1024-1024-1024-6 (the bundled agent's shape) on 5000 states: fp32 tier within 1e-5,
    bf16 tensor-core tier within 1e-3 of the output scale... of the fp32 oracle."""
    hidden = '1024-1024-1024'
    rs = np.random.RandomState(0)
    state = rs.normal(size=(5000, 615)).astype(np.float32)
    state[:, 315:] *= 0.3
    ac32, sd = _actor(hidden, 1111, 'fp32', kind)
    ac16, _ = _actor(hidden, 1111, 'bf16', kind)
    a_ref, _, pre_ref = O.actor_forward(sd, state, 0.0)
    st = torch.from_numpy(state).cuda()
    a32, _, pre32 = ac32.actor.forward_device(st, 0.0, want_pre=True)
    a16, _, pre16 = ac16.actor.forward_device(st, 0.0, want_pre=True)
    scale = np.abs(pre_ref).max()
    e32 = np.abs(pre32.cpu().numpy() - pre_ref).max() / scale
    e16 = np.abs(pre16.cpu().numpy() - pre_ref).max() / scale
    print('actor %s: fp32 tier rel err %.2e, bf16 tier rel err %.2e (scale %.3g)' % (kind, e32, e16, scale))
    assert e32 <= 1e-5
    assert e16 <= 1e-2
    assert np.abs(a32.cpu().numpy() - a_ref).max() <= 1e-5
    # row count taken from device memory, strided state rows (the env's 616-float pitch)
    buf = torch.zeros((5000, 616), device='cuda')
    buf[:, :615] = st
    n_dev = torch.tensor([1234], dtype=torch.int32, device='cuda')
    out = torch.full((5000, 3), 9.0, device='cuda')
    ac16.actor.forward_device(buf[:, :615], 0.0, n_rows_dev=n_dev, out_action=out, want_logp=False)
    assert torch.equal(out[:1234], a16[:1234]) and (out[1234:] == 9.0).all()


def test_actor_packed_bf16_state_matches_packing_path():
    """ttl_actor_forward_packed (bf16 rows supplied by the env) == ttl_actor_forward (packs fp32)."""
    ac16, _ = _actor('1024-1024-1024', 1111, 'bf16', 'tracking')
    rs = np.random.RandomState(1)
    st = torch.from_numpy(rs.normal(size=(3000, 615)).astype(np.float32)).cuda()
    sb = torch.zeros((3000, 640), dtype=torch.bfloat16, device='cuda')
    sb[:, :615] = st.to(torch.bfloat16)
    a1, _, p1 = ac16.actor.forward_device(st, 0.0, want_pre=True)
    a2, _, p2 = ac16.actor.forward_device(st, 0.0, want_pre=True, state_bf16=sb)
    assert torch.equal(p1, p2) and torch.equal(a1, a2)
