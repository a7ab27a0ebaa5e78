#!/usr/bin/env python
"""Runs bench.py's device-resident leg for a sequence of tiers in ONE process and prints, per leg, the
throughput and the per-kernel CUDA-event times: separates what a tier costs from where in the process
(power / clock state, allocation pattern) it happens to run.

    python benchmarks/tier_repeat.py fp16 fp16 bf16 bf16 fp16 tf32 fp16 [--steps 50] [--sleep 0.5]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('tiers', nargs='+')
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--sleep', type=float, default=0.0, help='idle seconds before every leg')
    a = ap.parse_args()
    a.operand_only, a.graph = True, False
    import torch
    from tracktolearn_b200 import _lib, synthetic
    dev = torch.device('cuda:0')
    lib = _lib.load()
    env, sub = B.make_env(B.SHAPE, B.VOXEL_MM, dev)
    env.seeds = B.sharded_seed_list(sub['seed_mask'].cpu().numpy(), 1, 0)
    sd = synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking')
    sampler = B.ClockSampler(0)
    sampler.start()

    def barrier():
        torch.cuda.synchronize(dev)
    for prec in a.tiers:
        if a.sleep:
            time.sleep(a.sleep)
        alg, r = B.run_tier(env, sd, prec, a, dev, lib, barrier, sampler, False)
        k = {n: round(1000.0 * ms / c, 1) for n, (c, ms) in sorted(r['prof'].items())}
        print('%-5s %6.1f M  %6.1f us/step  %s MHz  %s' % (prec, r['units'] / r['elapsed_ms'] / 1e3,
                                                          1000.0 * r['elapsed_ms'] / a.steps, r['sm_mhz'], k), flush=True)
        del alg
    sampler.stop()


if __name__ == '__main__':
    main()
