#!/usr/bin/env python
"""BASELINE.json configs[4]: TractOracle-Net batched scoring of finished streamlines.

    python benchmarks/oracle_bench.py [--n 1000000] [--cpu-n 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 benchmarks/oracle_bench.py        # N GPUs: contiguous chunks + all_gather of scores

Streamlines: ragged smooth random walks, lengths U{20..267} points; model:
n_head=4, n_layers=4, d=32, ff=2048, 128 tokens (assumed hyper-parameters, real blob missing).
Reports device-resident and host-to-host streamlines/s, achieved TFLOP/s against the
146.8 MFLOP/streamline figure, and the numpy oracle on the host cores for a small sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLOP_PER_STREAMLINE = 4 * 36700160 + 2 * 128 * 3 * 32 + 2 * 32


def make_streamlines(n, seed=0):
    rng = np.random.RandomState(seed)
    lens = rng.randint(20, 268, size=n)
    offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    total = int(offsets[-1])
    steps = rng.normal(size=(total, 3)).astype(np.float32)
    # smooth the directions a little and normalise to 0.75 voxel steps
    steps = steps + 2.0 * np.roll(steps, 1, axis=0) + np.roll(steps, 2, axis=0)
    steps *= 0.75 / np.linalg.norm(steps, axis=1, keepdims=True)
    pts = np.cumsum(steps, axis=0)
    starts = np.repeat(pts[offsets[:-1]], lens, axis=0)
    return (pts - starts + 40.0).astype(np.float32), offsets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=1000000)
    ap.add_argument('--cpu-n', type=int, default=256)
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'fp32'])
    a = ap.parse_args()
    import torch
    from tracktolearn_b200 import _lib, synthetic
    from tracktolearn_b200.oracles.oracle import OracleSingleton
    from tracktolearn_b200.tracking.tractogram import Tractogram
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    ck = synthetic.oracle_checkpoint(n_head=4, n_layers=4, input_size=384, seed=2222)
    model = OracleSingleton(ck, dev, precision=a.precision)
    data, offsets = make_streamlines(a.n)
    pts = torch.from_numpy(data).to(dev)
    off = torch.from_numpy(offsets).to(dev)
    model.predict_device(pts, off)          # warm-up (and NCCL set-up)
    barrier()
    lib = _lib.load()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scores = model.predict_device(pts, off)     # every rank its chunk + all_gather_into_tensor
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t[0]) * 1e-3
    lib.ttl_prof_enable(1)
    model.predict_device(pts, off)
    torch.cuda.synchronize()
    prof = _lib.prof_report()
    lib.ttl_prof_enable(0)
    barrier()
    t0 = time.perf_counter()
    host_scores = model.predict(Tractogram(data=data, offsets=offsets))
    barrier()
    host_s = time.perf_counter() - t0
    assert np.array_equal(host_scores, scores.cpu().numpy())
    if rank != 0:
        dist.destroy_process_group()
        return
    # CPU oracle on a sample
    from oracle import ttl_oracle as O
    sl = [data[offsets[i]:offsets[i + 1]] for i in range(a.cpu_n)]
    t0 = time.perf_counter()
    ref = O.oracle_predict(ck, sl)
    cpu_s = time.perf_counter() - t0
    err = float(np.abs(ref - host_scores[:a.cpu_n]).max())
    fwd_ms = sum(v[1] for k, v in prof.items() if k.startswith('oracle_forward'))
    n_rank = -(-a.n // world)
    out = {
        'metric': 'oracle streamlines/sec', 'n': a.n, 'n_gpus': world,
        'sharding': 'contiguous chunks of streamlines per rank, all_gather_into_tensor of the scores',
        'device_resident_streamlines_per_s': a.n / dev_s,
        'host_to_host_streamlines_per_s': a.n / host_s,
        'forward_kernel_tflops_rank0': n_rank * FLOP_PER_STREAMLINE / (fwd_ms * 1e-3) / 1e12 if fwd_ms else None,
        'kernels_ms': {k: v[1] for k, v in prof.items()},
        'cpu_port_streamlines_per_s': a.cpu_n / cpu_s, 'cpu_cores': os.cpu_count(), 'cpu_sample': a.cpu_n,
        'max_abs_err_vs_cpu_oracle': err,
        'dtype': ('f16 operands on tcgen05 (QKV, QK^T, PV, out-proj, FFN), f32 accumulate / softmax / LayerNorm'
                  if a.precision == 'fp16' else 'f32 (CUDA cores)'),
        'flop_per_streamline': FLOP_PER_STREAMLINE,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
