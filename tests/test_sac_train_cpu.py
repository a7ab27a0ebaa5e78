"""SAC-auto update against a fixture recorded from the reference's SACAuto.update, the device
replay buffer's ring semantics, and the data-parallel gradient all-reduce (gloo, world size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_golden
from tracktolearn_b200.algorithms.sac_train import SACAutoLearner
from tracktolearn_b200.algorithms.shared.replay import OffPolicyReplayBuffer


def _learner_from(g):
    L = SACAutoLearner(24, 3, '16-12', lr=3e-4, gamma=0.95, alpha=0.2, device='cpu')
    L.actor.load_state_dict({k[len('actor0.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('actor0.')})
    L.critic.load_state_dict({k[len('critic0.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('critic0.')})
    L.target_actor.load_state_dict(L.actor.state_dict())
    L.target_critic.load_state_dict(L.critic.state_dict())
    return L


def _batch(g):
    return tuple(torch.from_numpy(g['batch.' + n]) for n in ('state', 'action', 'next_state', 'reward', 'not_done'))


def test_update_matches_reference_fixture():
    g = load_golden('sac_update')
    L = _learner_from(g)
    torch.manual_seed(77)
    for _ in range(2):
        L.update(_batch(g))
    for k, v in L.actor.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g['actor2.' + k], rtol=1e-5, atol=1e-7)
    for k, v in L.critic.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g['critic2.' + k], rtol=1e-5, atol=1e-7)
    for k, v in L.target_critic.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g['target_critic2.' + k], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(L.log_alpha.detach().numpy(), g['log_alpha2'], rtol=1e-6)


def test_replay_buffer_ring_and_sampling():
    rb = OffPolicyReplayBuffer(5, 3, max_size=10, device='cpu')
    s = torch.arange(35, dtype=torch.float32).reshape(7, 5)
    rb.add(s, torch.zeros(7, 3), s + 1, torch.arange(7.), torch.tensor([0, 0, 1, 0, 0, 0, 1.]))
    rb.add(s + 100, torch.ones(7, 3), s + 101, torch.arange(7.) + 10, torch.zeros(7))
    assert len(rb) == 10 and rb.ptr == 4
    np.testing.assert_array_equal(rb.state[:4].numpy(), (s + 100)[3:].numpy())      # wrapped around
    np.testing.assert_array_equal(rb.state[7:].numpy(), (s + 100)[:3].numpy())
    np.testing.assert_array_equal(rb.not_done[4:7, 0].numpy(), [1, 1, 0])
    st, a, ns, r, nd = rb.sample(6, generator=torch.Generator().manual_seed(0))
    assert st.shape == (6, 5) and len(set(r.tolist())) == 6                         # without replacement
    np.testing.assert_array_equal(ns.numpy(), st.numpy() + 1)


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = load_golden('sac_update')
    L = _learner_from(g)
    batch = _batch(g)
    half = tuple(t[rank * 8:(rank + 1) * 8] for t in batch)          # each rank sees half the batch
    eps = torch.Generator().manual_seed(9)
    e1, e2 = torch.randn((16, 3), generator=eps), torch.randn((16, 3), generator=eps)
    L.update(half, eps=e1[rank * 8:(rank + 1) * 8], eps_next=e2[rank * 8:(rank + 1) * 8])
    if rank == 0:
        q.put({k: v.numpy().copy() for k, v in L.actor.state_dict().items()})
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_update_equals_full_batch_update():
    """Mean-of-rank-gradients == gradient of the full batch: two ranks with half a batch each end
    up with the weights a single process gets from the whole batch."""
    g = load_golden('sac_update')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L = _learner_from(g)
    eps = torch.Generator().manual_seed(9)
    e1, e2 = torch.randn((16, 3), generator=eps), torch.randn((16, 3), generator=eps)
    L.update(_batch(g), eps=e1, eps_next=e2)
    for k, v in L.actor.state_dict().items():
        np.testing.assert_allclose(got[k], v.numpy(), rtol=2e-5, atol=1e-7)
