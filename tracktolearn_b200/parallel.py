"""Multi-GPU plumbing for tracking: seed sharding and the final tractogram gather.

The path shards by seed: every rank holds a replica of the volume,
mask and actor, tracks its own contiguous slice of the (globally shuffled) seeds and there is
no collective on the data path.  The only exchange is at the end: ranks send their packed
streamlines to rank 0, which concatenates them in rank order -- the same order a single GPU
would have produced.  Works with any ``torch.distributed`` backend (NCCL on GPUs, gloo in the
CPU tests): only host logic lives here.
"""
import numpy as np
import torch
import torch.distributed as dist

from tracktolearn_b200.tracking.tractogram import Tractogram


def shard_bounds(n, rank, world):
    """Contiguous slice [start, end) of n items for `rank`; sizes differ by at most one and
    concatenating the slices in rank order gives back 0..n."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_seeds(seeds, rank=None, world=None):
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    s, e = shard_bounds(len(seeds), rank, world)
    return seeds[s:e]


def sharded_scores(n_items, score_chunk, device=None):
    """Per-item scores computed by all ranks together (oracle scoring): rank k
    scores the contiguous chunk ``shard_bounds(n_items, k, world)`` with ``score_chunk(start, end) ->
    float32 tensor [end - start]`` and one ``all_gather_into_tensor`` of the (padded) chunks hands every
    rank the full ``[n_items]`` tensor, in item order.  Without torch.distributed: ``score_chunk(0, n)``."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return score_chunk(0, n_items)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else _device_for_backend()
    s, e = shard_bounds(n_items, rank, world)
    pad = -(-n_items // world)                       # chunk sizes differ by at most one
    mine = torch.zeros((pad,), dtype=torch.float32, device=dev)
    if e > s:
        mine[:e - s] = score_chunk(s, e).to(dev, dtype=torch.float32).reshape(-1)
    full = torch.empty((world * pad,), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(full, mine)
    parts = []
    for r in range(world):
        rs, re = shard_bounds(n_items, r, world)
        parts.append(full[r * pad:r * pad + (re - rs)])
    return torch.cat(parts) if parts else full[:0]


def _device_for_backend():
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


_PINNED = {}


def _pinned(nbytes):
    """Grow-only pinned staging buffer (a fresh 1 GB cudaHostAlloc per gather costs more than the
    transfer it serves)."""
    cur = _PINNED.get('buf')
    if cur is None or cur.numel() < nbytes:
        cur = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8, pin_memory=True)
        _PINNED['buf'] = cur
    return cur


def _pinned_named(name, nbytes):
    cur = _PINNED.get(name)
    if cur is None or cur.numel() < nbytes:
        cur = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        _PINNED[name] = cur
    return cur


def gather_packed(points, lengths, seeds, flags, dst=0, copy=True):
    """Gather packed streamlines on rank `dst` in rank order.

    ``points`` [n_pts, 3] float32, ``lengths`` [n] int64, ``seeds`` [n, 3] float64, ``flags`` [n] int64:
    tensors on this rank's device for the backend (CUDA under NCCL -- e.g. straight from
    ``env.get_streamlines_device()``, no host round trip -- CPU under gloo).  One all_gather of the
    sizes, then every rank sends exactly its rows into its slice of `dst`'s buffers (batched
    point-to-point: device-to-device over NVLink under NCCL) and `dst` makes one D2H copy per array
    into grow-only pinned memory.  With ``copy=False`` the returned arrays are views of that pinned
    memory, valid until the next gather.  Returns a Tractogram on `dst`, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device_for_backend()
    parts = [(points.to(dev, dtype=torch.float32).reshape(-1, 3).contiguous(), 3, torch.float32, 'pts'),
             (lengths.to(dev, dtype=torch.int64).reshape(-1, 1).contiguous(), 1, torch.int64, 'len'),
             (seeds.to(dev, dtype=torch.float64).reshape(-1, 3).contiguous(), 3, torch.float64, 'seeds'),
             (flags.to(dev, dtype=torch.int64).reshape(-1, 1).contiguous(), 1, torch.int64, 'flags')]
    n_sl, n_pts = int(parts[1][0].shape[0]), int(parts[0][0].shape[0])
    sizes = torch.tensor([n_sl, n_pts], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    n_per, p_per = all_sizes[:, 0], all_sizes[:, 1]
    ops, bufs = [], []
    for mine, width, dtype, name in parts:
        counts = p_per if name == 'pts' else n_per
        if rank == dst:
            buf = torch.empty((int(counts.sum()), width), dtype=dtype, device=dev)
            o = 0
            for r in range(world):
                c = int(counts[r])
                if r == dst:
                    buf[o:o + c].copy_(mine)
                elif c > 0:
                    ops.append(dist.P2POp(dist.irecv, buf[o:o + c], r))
                o += c
            bufs.append(buf)
        elif int(counts[rank]) > 0:
            ops.append(dist.P2POp(dist.isend, mine, dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    host = []
    for buf, (_, _, dtype, name) in zip(bufs, parts):
        if buf.is_cuda:
            nbytes = buf.numel() * buf.element_size()
            h = _pinned_named(name, nbytes)[:nbytes].view(dtype).view(buf.shape)
            h.copy_(buf, non_blocking=True)
            host.append(h)
        else:
            host.append(buf)
    if bufs and bufs[0].is_cuda:
        torch.cuda.current_stream(bufs[0].device).synchronize()
    arrs = [h.numpy().copy() if copy else h.numpy() for h in host]
    data, lens, gseeds, gflags = arrs
    offsets = np.concatenate(([0], np.cumsum(lens[:, 0]))).astype(np.int64)
    return Tractogram(data=data, offsets=offsets, data_per_streamline={'seeds': gseeds, 'flags': gflags[:, 0]})


def gather_tractogram(local, dst=0, copy=True):
    """Gather host-side packed tractograms (``Tractogram``) on rank `dst` in rank order; returns the
    merged Tractogram on `dst` and None elsewhere.  Prefer ``gather_env_streamlines`` when the
    streamlines are still on the device."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    n_sl = len(local)
    return gather_packed(
        torch.as_tensor(np.ascontiguousarray(local.data, dtype=np.float32)).reshape(-1, 3),
        torch.as_tensor(np.diff(local.offsets).astype(np.int64)),
        torch.as_tensor(np.ascontiguousarray(local.data_per_streamline.get('seeds', np.zeros((n_sl, 3))),
                                             dtype=np.float64)).reshape(-1, 3),
        torch.as_tensor(np.asarray(local.data_per_streamline.get('flags', np.zeros(n_sl))).astype(np.int64)),
        dst=dst, copy=copy)


def gather_env_streamlines(env, dst=0, copy=True):
    """The final exchange of a multi-GPU tracking run, from the device: this rank's packed streamlines
    (``env.get_streamlines_device()``), seeds and flags go to rank `dst` without touching the host on
    the sending side.  Single process: the env's own tractogram."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return env.get_streamlines(copy=copy)
    pts, offsets = env.get_streamlines_device()
    n = int(offsets.shape[0]) - 1
    seeds = torch.as_tensor(np.ascontiguousarray(env.initial_points, dtype=np.float64)).reshape(-1, 3)
    return gather_packed(pts, offsets[1:] - offsets[:-1], seeds, env._batch.flags[:n].to(torch.int64),
                         dst=dst, copy=copy)
