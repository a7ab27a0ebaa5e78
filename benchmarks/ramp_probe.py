#!/usr/bin/env python
"""How long does a freshly leased, idle B200 take to reach its steady boost state?  One process, one env and
actor; every `--period` seconds: 16 untimed + 20 timed steps of the configs[1] step (fp16 tier), printed with the
SM clock / power NVML reports right after.  `--busy` keeps a light kernel running between the samples.

    python benchmarks/ramp_probe.py [--seconds 90] [--period 3] [--busy]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=40.0)   # 873 k seeds last ~900 steps after the burn-in
    ap.add_argument('--period', type=float, default=3.0)
    ap.add_argument('--busy', action='store_true')
    a = ap.parse_args()
    import torch
    import pynvml
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.algorithms.rl import StepRunner
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device('cuda:0')
    t_first = time.monotonic()
    env, sub = B.make_env(B.SHAPE, B.VOXEL_MM, dev)
    env.seeds = B.sharded_seed_list(sub['seed_mask'].cpu().numpy(), 1, 0)
    alg = SACAuto(B.STATE_SIZE, 3, B.HIDDEN, n_actors=B.N_ACTOR, device=dev, precision='fp16')
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking'))
    env.reset_streaming(0, len(env.seeds), B.N_ACTOR, fp32_state=False, operand='fp16')
    runner = StepRunner(env, alg.agent.actor, 0.0, use_graph=False)
    stream = torch.cuda.current_stream(dev)
    for _ in range(384):
        runner.step()
    torch.cuda.synchronize(dev)
    filler = torch.zeros((1 << 20,), device=dev)
    out = []
    t0 = time.monotonic()
    while time.monotonic() - t0 < a.seconds:
        for _ in range(16):
            runner.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            runner.step()
        e1.record(stream)
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        torch.cuda.synchronize(dev)
        us = 1000.0 * e0.elapsed_time(e1) / 20
        rec = {'t_since_setup': round(time.monotonic() - t_first, 1), 'us_per_step': round(us, 1), 'sm_mhz': clk, 'watts': round(pw)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        t_next = time.monotonic() + a.period
        while time.monotonic() < t_next:
            if a.busy:
                filler.add_(1.0)
                torch.cuda.synchronize(dev)
            else:
                time.sleep(0.05)


if __name__ == '__main__':
    main()
