"""world_size-2 gloo test of the multi-GPU host logic: seed shards tile the seed list in rank
order and the gathered tractogram equals the single-process one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tracktolearn_b200 import parallel
from tracktolearn_b200.tracking.tractogram import Tractogram


def _fake_track(seeds):
    """Deterministic per-seed 'streamlines' (length and points depend only on the seed)."""
    sl = []
    for s in seeds:
        L = 2 + int(abs(s[0] * 7 + s[1] * 3)) % 9
        sl.append((s[None, :] + 0.5 * np.arange(L)[:, None]).astype(np.float32))
    return Tractogram(streamlines=sl, data_per_streamline={
        'seeds': np.asarray(seeds, dtype=np.float64).reshape(-1, 3),
        'flags': (np.arange(len(seeds)) % 7).astype(np.int64)})


def _worker(rank, world, port, seeds, q, via='nccl'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = parallel.shard_seeds(seeds)
    merged = parallel.gather_tractogram(_fake_track(mine), via=via)
    if rank == 0:
        q.put((merged.data, merged.offsets, merged.data_per_streamline['seeds'],
               merged.data_per_streamline['flags'], len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_tile_the_range():
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            got = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in got]
            assert max(sizes) - min(sizes) <= 1


import pytest


@pytest.mark.parametrize('via', ['nccl', 'host'])
def test_two_rank_gather_equals_single_process(via):
    """'nccl': point-to-point into rank 0's buffers; 'host': every rank copies into a shared host arena
    (the single-node path; here over gloo on CPU, without the page-locking)."""
    rs = np.random.RandomState(0)
    seeds = rs.uniform(0, 20, size=(101, 3))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seeds, q, via)) for r in range(2)]
    for p in procs:
        p.start()
    data, offsets, gseeds, gflags, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref0 = _fake_track(seeds[:n0])
    ref1 = _fake_track(seeds[n0:])
    np.testing.assert_array_equal(data, np.concatenate((ref0.data, ref1.data)))
    np.testing.assert_array_equal(np.diff(offsets), np.concatenate((ref0.lengths, ref1.lengths)))
    np.testing.assert_array_equal(gseeds, seeds)
    np.testing.assert_array_equal(gflags, np.concatenate((np.arange(n0) % 7, np.arange(101 - n0) % 7)))
    assert n0 == 51


def _score_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    calls = []

    def score_chunk(s, e):                      # stands in for TractOracle-Net: a function of the item index
        calls.append((s, e))
        return torch.arange(s, e, dtype=torch.float32) * 0.5 + 1.0
    full = parallel.sharded_scores(n, score_chunk)
    q.put((rank, full.numpy(), calls))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_scores_cover_every_item_once():
    """contiguous chunks of streamlines per rank, all_gather_into_tensor of the scores;
    every rank ends with all N scores in item order, each item scored by exactly one rank."""
    for n in (1, 2, 101):
        with socket.socket() as s:
            s.bind(('127.0.0.1', 0))
            port = s.getsockname()[1]
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_score_worker, args=(r, 2, port, n, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in range(2)]
        for p in procs:
            p.join(timeout=60)
        want = np.arange(n, dtype=np.float32) * 0.5 + 1.0
        covered = []
        for rank, full, calls in got:
            np.testing.assert_array_equal(full, want)
            covered += [i for (s, e) in calls for i in range(s, e)]
        assert sorted(covered) == list(range(n))
    # single process: one call over everything
    out = parallel.sharded_scores(5, lambda s, e: torch.arange(s, e, dtype=torch.float32))
    np.testing.assert_array_equal(out.numpy(), np.arange(5, dtype=np.float32))
