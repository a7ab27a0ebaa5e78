#!/usr/bin/env python
"""How long does one step take when the GPU work is tiny (256 rows)?  Wall clock per step of the
device loop (StepRunner) vs the CUDA-event time of the same steps: the difference is host launch
overhead, which bounds the low-occupancy tail of an episode."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests.test_tracker_gpu import _setup
    from tracktolearn_b200.algorithms.rl import StepRunner
    env, alg, sub, seeds, sd = _setup(shape=(48, 52, 44), n_seeds=60000, precision='bf16')
    hidden = os.environ.get('PROBE_HIDDEN', '1024-1024-1024')
    rows = int(os.environ.get('PROBE_ROWS', '256'))
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    alg = SACAuto(615, 3, hidden, n_actors=rows, device=torch.device('cuda:0'), precision='bf16')
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(615, hidden, seed=5, kind='tracking'))
    out = {}
    for graph in (False, True):
        env.reset_streaming(0, len(seeds), rows, fp32_state=False)
        r = StepRunner(env, alg.agent.actor, 0.0, use_graph=graph)
        for _ in range(50):
            r.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 1000
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            r.step()
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        out['graph' if graph else 'plain'] = {'host_issue_us_per_step': 1e6 * t_issue / n,
                                              'wall_us_per_step': 1e6 * t_all / n,
                                              'device_us_per_step': 1e3 * e0.elapsed_time(e1) / n,
                                              'alive': env.n_alive()}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
