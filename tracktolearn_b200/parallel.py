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


def gather_tractogram(local, dst=0):
    """Gather packed tractograms on rank `dst` in rank order.  Returns the merged Tractogram on
    `dst` and None elsewhere.  Three variable-size all-gathers worth of data, done as one
    all_gather of the sizes followed by padded all_gathers (NCCL has no gatherv)."""
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
    max_sl, max_pts = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())

    def padded_gather(arr, rows, width, dtype):
        buf = torch.zeros((rows, width), dtype=dtype, device=dev)
        if len(arr):
            buf[:len(arr)] = torch.as_tensor(np.ascontiguousarray(arr).reshape(len(arr), width)).to(dev, dtype=dtype)
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        return [o.cpu().numpy() for o in out] if rank == dst else None

    pts = padded_gather(local.data, max_pts, 3, torch.float32)
    lens = padded_gather(np.diff(local.offsets).astype(np.int64), max_sl, 1, torch.int64)
    seeds = padded_gather(np.asarray(local.data_per_streamline.get('seeds', np.zeros((n_sl, 3)))),
                          max_sl, 3, torch.float64)
    flags = padded_gather(np.asarray(local.data_per_streamline.get('flags', np.zeros(n_sl))).astype(np.int64),
                          max_sl, 1, torch.int64)
    if rank != dst:
        return None
    data = np.concatenate([pts[r][:all_sizes[r, 1]] for r in range(world)])
    all_lens = np.concatenate([lens[r][:all_sizes[r, 0], 0] for r in range(world)])
    offsets = np.concatenate(([0], np.cumsum(all_lens))).astype(np.int64)
    return Tractogram(
        data=data, offsets=offsets,
        data_per_streamline={
            'seeds': np.concatenate([seeds[r][:all_sizes[r, 0]] for r in range(world)]),
            'flags': np.concatenate([flags[r][:all_sizes[r, 0], 0] for r in range(world)])})
