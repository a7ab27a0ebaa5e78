#!/usr/bin/env python
"""Step time against the number of alive streamlines (the low-occupancy tail of a tracking run): a
streaming batch of `rows` slots over plenty of seeds keeps exactly `rows` streamlines alive; 40 timed
steps each, per-kernel CUDA-event times.

    python benchmarks/tail_probe.py [--precision fp16] [--rows 50000 20000 8000 4000 2000 1000 256 32]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--rows', type=int, nargs='+', default=[50000, 20000, 8000, 4000, 2000, 1000, 256, 32])
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--graph', action='store_true')
    a = ap.parse_args()
    import torch
    from tracktolearn_b200 import _lib, synthetic
    from tracktolearn_b200.algorithms.rl import StepRunner
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    dev = torch.device('cuda:0')
    lib = _lib.load()
    env, sub = B.make_env(B.SHAPE, B.VOXEL_MM, dev)
    env.seeds = B.sharded_seed_list(sub['seed_mask'].cpu().numpy(), 1, 0)
    alg = SACAuto(B.STATE_SIZE, 3, B.HIDDEN, n_actors=B.N_ACTOR, device=dev, precision=a.precision)
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking'))
    stream = torch.cuda.current_stream(dev)
    out = {}
    for rows in a.rows:
        env.reset_streaming(0, len(env.seeds), rows, fp32_state=False, operand=a.precision)
        runner = StepRunner(env, alg.agent.actor, 0.0, use_graph=a.graph)
        for _ in range(70):
            runner.step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            runner.step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        us = 1000.0 * e0.elapsed_time(e1) / a.steps
        lib.ttl_prof_enable(1)
        plain = StepRunner(env, alg.agent.actor, 0.0, use_graph=False)
        for _ in range(10):
            plain.step()
        torch.cuda.synchronize(dev)
        prof = _lib.prof_report()
        lib.ttl_prof_enable(0)
        k = {n: round(1000.0 * ms / c, 1) for n, (c, ms) in sorted(prof.items())}
        print('rows %6d  %7.1f us/step  %s' % (rows, us, k), flush=True)
        out[rows] = {'us_per_step': us, 'kernels_us': k}
    return out


if __name__ == '__main__':
    main()
