"""Device TractOracle-Net (resample + diff + transformer) against the reference-recorded fixture
and the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O
from tests.helpers import load_golden, split_by_counts
from tracktolearn_b200 import synthetic

pytestmark = pytest.mark.gpu


# fp16 tensor-core tier: linear1 / linear2 take fp16 operands (what the reference's CUDA path does
# under autocast); scores are sigmoid outputs in [0, 1], tolerance absolute.
FP16_TOL = 5e-3


def _oracle(ck, precision='fp32'):
    from tracktolearn_b200.oracles.oracle import OracleSingleton
    OracleSingleton.clear()
    return OracleSingleton(ck, torch.device('cuda:0'), batch_size=4096, precision=precision)


def test_oracle_net_matches_reference_fixture():
    g = load_golden('oracle_net')
    n_head, n_layers, input_size, seed = [int(v) for v in g['hp']]
    ck = synthetic.oracle_checkpoint(n_head=n_head, n_layers=n_layers, input_size=input_size, seed=seed)
    sl = split_by_counts(g['sl_points'], g['sl_lengths'])
    model = _oracle(ck)
    scores = model.predict(sl)
    np.testing.assert_allclose(scores, g['scores'], rtol=0, atol=2e-5)
    # front end alone: resample(128) + diff, bit-level agreement expected up to float rounding
    from tracktolearn_b200 import _lib
    pts = torch.from_numpy(g['sl_points']).cuda()
    off = torch.from_numpy(np.concatenate(([0], np.cumsum(g['sl_lengths']))).astype(np.int64)).cuda()
    dirs = torch.empty((len(sl), 127, 3), device='cuda')
    _lib.check(_lib.load().ttl_oracle_features(_lib.ptr(pts), _lib.ptr(off), len(sl), _lib.ptr(dirs),
                                               _lib.stream_ptr(torch.device('cuda:0'))), 'features')
    np.testing.assert_allclose(dirs.cpu().numpy(), g['dirs'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('n_head,n_layers', [(4, 4), (2, 1), (8, 2)])
def test_oracle_net_matches_numpy_oracle(n_head, n_layers):
    ck = synthetic.oracle_checkpoint(n_head=n_head, n_layers=n_layers, seed=77)
    rng = np.random.RandomState(n_head * 10 + n_layers)
    sl = synthetic.random_streamlines(300, rng, min_pts=2, max_pts=200)
    sl[0] = sl[0][:1]                       # single point
    sl[1] = np.repeat(sl[1][:1], 5, axis=0)  # zero-length
    model = _oracle(ck)
    model.batch_size = 128                   # exercise chunking (300 = 2*128 + 44)
    got = model.predict(sl)
    ref = O.oracle_predict(ck, sl)
    np.testing.assert_allclose(got, ref, rtol=0, atol=5e-5)


def test_oracle_net_fp16_tensor_core_tier_matches_reference_fixture():
    g = load_golden('oracle_net')
    n_head, n_layers, input_size, seed = [int(v) for v in g['hp']]
    ck = synthetic.oracle_checkpoint(n_head=n_head, n_layers=n_layers, input_size=input_size, seed=seed)
    sl = split_by_counts(g['sl_points'], g['sl_lengths'])
    scores = _oracle(ck, 'fp16').predict(sl)
    np.testing.assert_allclose(scores, g['scores'], rtol=0, atol=FP16_TOL)


@pytest.mark.parametrize('n_head,n_layers,n', [(4, 4, 700), (2, 1, 1), (8, 2, 300)])
def test_oracle_net_fp16_tensor_core_tier_matches_numpy_oracle(n_head, n_layers, n):
    """More streamlines than resident CTAs (2 x 148) so the persistent loop, the weight ring and every
    barrier wrap around across streamlines; n = 1 covers a single CTA."""
    ck = synthetic.oracle_checkpoint(n_head=n_head, n_layers=n_layers, seed=78)
    rng = np.random.RandomState(n_head * 10 + n_layers)
    sl = synthetic.random_streamlines(n, rng, min_pts=2, max_pts=200)
    model = _oracle(ck, 'fp16')
    got = model.predict(sl)
    ref = O.oracle_predict(ck, sl)
    np.testing.assert_allclose(got, ref, rtol=0, atol=FP16_TOL)
    # same inputs, fp32 tier: the two tiers agree on which side of 0.5 every clear-cut score lies
    got32 = _oracle(ck, 'fp32').predict(sl)
    clear = np.abs(got32 - 0.5) > 2 * FP16_TOL
    np.testing.assert_array_equal(got[clear] > 0.5, got32[clear] > 0.5)
    # deterministic
    np.testing.assert_array_equal(_oracle(ck, 'fp16').predict(sl), got)
