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


def gather_tractogram(local, dst=0):
    """Gather packed tractograms on rank `dst` in rank order.  Returns the merged Tractogram on
    `dst` and None elsewhere.  One all_gather of the sizes, then every rank sends exactly its rows
    straight into its slice of rank `dst`'s buffer (batched point-to-point: NVLink device-to-device
    under NCCL, one pinned D2H on `dst`; plain CPU tensors under gloo)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device_for_backend()
    n_sl = len(local)
    n_pts = int(local.offsets[-1]) if n_sl else 0
    sizes = torch.tensor([n_sl, n_pts], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()

    def gather_rows(arr, counts, width, dtype):
        mine = torch.as_tensor(np.ascontiguousarray(arr).reshape(-1, width)).to(dev, dtype=dtype).contiguous()
        ops, buf = [], None
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
        elif int(counts[rank]) > 0:
            ops.append(dist.P2POp(dist.isend, mine, dst))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if rank != dst:
            return None
        if buf.is_cuda:
            host = _pinned(buf.numel() * buf.element_size())[:buf.numel() * buf.element_size()].view(dtype)
            host = host.view(buf.shape)
            host.copy_(buf)
            return host.numpy().copy()
        return buf.numpy()

    n_per, p_per = all_sizes[:, 0], all_sizes[:, 1]
    data = gather_rows(local.data, p_per, 3, torch.float32)
    lens = gather_rows(np.diff(local.offsets).astype(np.int64), n_per, 1, torch.int64)
    seeds = gather_rows(np.asarray(local.data_per_streamline.get('seeds', np.zeros((n_sl, 3)))), n_per, 3,
                        torch.float64)
    flags = gather_rows(np.asarray(local.data_per_streamline.get('flags', np.zeros(n_sl))).astype(np.int64),
                        n_per, 1, torch.int64)
    if rank != dst:
        return None
    offsets = np.concatenate(([0], np.cumsum(lens[:, 0]))).astype(np.int64)
    return Tractogram(data=data, offsets=offsets,
                      data_per_streamline={'seeds': seeds, 'flags': flags[:, 0]})
