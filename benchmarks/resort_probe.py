#!/usr/bin/env python
"""north_star (1), "voxel-local streamlines kept together by a per-step tip sort": steady-state step time
of BASELINE configs[1] (50 000 slots) with the alive list re-sorted by tip voxel every k steps
(ttl_env_resort), k = 0 (never: the reset-time voxel-raster slot order only), 32, 16, 8, 4, 1.

    python benchmarks/resort_probe.py [--precision fp16] [--every 0 32 16 8 4 1] [--steps 192]

Prints the mean step time INCLUDING the amortised sorts and the per-kernel CUDA-event times."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--every', type=int, nargs='+', default=[0, 32, 16, 8, 4, 1])
    ap.add_argument('--steps', type=int, default=192)
    ap.add_argument('--rows', type=int, default=B.N_ACTOR)
    ap.add_argument('--no-locality', action='store_true', help='seeds take slots in shuffled row order')
    ap.add_argument('--hold', type=int, default=0, help='run this many extra steps of the LAST setting (for ncu)')
    a = ap.parse_args()
    import torch
    from tracktolearn_b200 import _lib, synthetic
    from tracktolearn_b200.algorithms.rl import StepRunner
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    dev = torch.device('cuda:0')
    lib = _lib.load()
    env, sub = B.make_env(B.SHAPE, B.VOXEL_MM, dev)
    env.seeds = B.sharded_seed_list(sub['seed_mask'].cpu().numpy(), 1, 0)
    alg = SACAuto(B.STATE_SIZE, 3, B.HIDDEN, n_actors=a.rows, device=dev, precision=a.precision)
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking'))
    stream = torch.cuda.current_stream(dev)
    for every in a.every:
        env.reset_streaming(0, len(env.seeds), a.rows, fp32_state=False, operand=a.precision, locality=not a.no_locality)
        runner = StepRunner(env, alg.agent.actor, 0.0, use_graph=False)

        def run(n, it0):
            for it in range(it0, it0 + n):
                if every and it % every == 0:
                    env.resort_device()
                runner.step()
            return it0 + n
        it = run(B.BURN_IN, 1)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        it = run(a.steps, it)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        us = 1000.0 * e0.elapsed_time(e1) / a.steps
        lib.ttl_prof_enable(1)
        it = run(32, it)
        torch.cuda.synchronize(dev)
        prof = _lib.prof_report()
        lib.ttl_prof_enable(0)
        k = {n: round(1000.0 * ms / c, 1) for n, (c, ms) in sorted(prof.items())}
        print('resort every %3d  %7.1f us/step  %s' % (every, us, k), flush=True)
    if a.hold:
        run(a.hold, it)
        torch.cuda.synchronize(dev)


if __name__ == '__main__':
    main()
