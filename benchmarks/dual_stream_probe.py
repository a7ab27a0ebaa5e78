#!/usr/bin/env python
"""Experiment: does running two half-size batches out of phase on two CUDA streams (one batch in
the tensor-bound actor while the other is in the latency/L2-bound env kernels) beat one full-size
batch?  Prints aggregate streamline-steps/s for both arrangements on the bench workload."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def make(dev, sub, n_actor, rank_seed):
    import torch
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.algorithms.rl import StepRunner
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import NoisyTrackingEnvironment
    affine = np.diag([B.VOXEL_MM] * 3 + [1.0])
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['seed_mask'], affine), None, affine)
    dto = {'n_dirs': 100, 'theta': B.THETA, 'npv': 1, 'binary_stopping_threshold': 0.1,
           'step_size': B.STEP_MM, 'min_length': 10.0, 'max_length': B.MAX_LENGTH_MM,
           'oracle_checkpoint': None, 'oracle_stopping_criterion': False, 'scoring_data': None,
           'compute_reward': False, 'alignment_weighting': 0.0, 'oracle_bonus': 0.0,
           'rng': np.random.RandomState(1337), 'device': dev, 'target_sh_order': 8,
           'noise': 0.0, 'fa_map': None, 'state_of_stopped': False}
    env = NoisyTrackingEnvironment(subject, 'testing', dto)
    env.seeds = B.draw_seeds(sub['seed_mask'].cpu().numpy(), rank_seed)
    alg = SACAuto(B.STATE_SIZE, 3, B.HIDDEN, n_actors=n_actor, device=dev, precision='bf16')
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking'))
    env.reset_streaming(0, len(env.seeds), n_actor, fp32_state=False)
    return env, alg, StepRunner(env, alg.agent.actor, 0.0, use_graph=False)


def run(dev, parts, steps, streams):
    import torch
    for _ in range(B.BURN_IN + 20):
        for (env, alg, r), s in zip(parts, streams):
            with torch.cuda.stream(s):
                r.step()
    torch.cuda.synchronize(dev)
    for env, _, _ in parts:
        env.n_alive()
    before = sum(env.streamline_steps() for env, _, _ in parts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    for s in streams:
        s.wait_stream(torch.cuda.current_stream(dev))
    for _ in range(steps):
        for (env, alg, r), s in zip(parts, streams):
            with torch.cuda.stream(s):
                r.step()
    for s in streams:
        torch.cuda.current_stream(dev).wait_stream(s)
    e1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    for env, _, _ in parts:
        env.n_alive()
    units = sum(env.streamline_steps() for env, _, _ in parts) - before
    ms = e0.elapsed_time(e1)
    return units / (ms * 1e-3), ms / steps


def main():
    import torch
    from tracktolearn_b200 import synthetic
    dev = torch.device('cuda:0')
    sub = synthetic.make_subject(B.SHAPE, seed=1234, device=dev, with_peaks=False)
    out = {}
    one = [make(dev, sub, B.N_ACTOR, 0)]
    v, ms = run(dev, one, 200, [torch.cuda.current_stream(dev)])
    out['one_batch_50000'] = {'steps_per_s': v, 'ms_per_round': ms}
    del one
    torch.cuda.empty_cache()
    for n_parts in (2, 3):
        parts = [make(dev, sub, B.N_ACTOR // n_parts, i) for i in range(n_parts)]
        streams = [torch.cuda.Stream(dev) for _ in range(n_parts)]
        v, ms = run(dev, parts, 200, streams)
        out['%d_batches_out_of_phase' % n_parts] = {'steps_per_s': v, 'ms_per_round': ms}
        del parts
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
