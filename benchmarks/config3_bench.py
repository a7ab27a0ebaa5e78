#!/usr/bin/env python
"""BASELINE.json configs[2]: 0.5 mm-iso synthetic 290^3 order-8 fODF (auto step 0.375 mm, 798 steps
max), 1 000 000 seeds sharded across the GPUs of one box (strong scaling: the total is fixed).

    python benchmarks/config3_bench.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 benchmarks/config3_bench.py       # N GPUs, one rank per GPU

Every rank holds a replica of the volume (4.68 GB channel-padded), the mask coefficients and the
actor, tracks its contiguous slice of the shuffled seeds with the streaming tracker (n_actor
50 000 slots per GPU) and hands its packed streamlines to rank 0 at the end (NCCL; the only
collective).  Timed on the device (CUDA events), max over ranks; the tractogram gather is timed
separately.  Also checks size-independent properties of the result at full size (every seed yields
one streamline that starts on its seed, lengths within [2, max_nb_steps + 1], flags non-zero)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

SHAPE = (290, 290, 290)
VOXEL_MM = 0.5
N_SEEDS = 1000000
STEP_MM = VOXEL_MM / B.TRAINED_VOXEL * B.TRAINED_STEP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=N_SEEDS)
    ap.add_argument('--n-actor', type=int, default=B.N_ACTOR)
    ap.add_argument('--shape', type=int, nargs=3, default=list(SHAPE))
    ap.add_argument('--no-gather', action='store_true')
    ap.add_argument('--graph', action='store_true', help='replay the step from CUDA graphs')
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'tf32', 'bf16'])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tracktolearn_b200 import parallel, synthetic
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import NoisyTrackingEnvironment
    from tracktolearn_b200.tracking.tracker import Tracker
    from tracktolearn_b200.tracking.tractogram import Tractogram

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    shape = tuple(a.shape)
    t_setup = time.perf_counter()
    sub = synthetic.make_subject(shape, seed=1234, device=dev, with_peaks=False)
    affine = np.diag([VOXEL_MM] * 3 + [1.0])
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['seed_mask'], affine), None, affine)
    dto = {'n_dirs': 100, 'theta': B.THETA, 'npv': 1, 'binary_stopping_threshold': 0.1,
           'step_size': STEP_MM, 'min_length': 10.0, 'max_length': B.MAX_LENGTH_MM,
           'oracle_checkpoint': None, 'oracle_stopping_criterion': False, 'scoring_data': None,
           'compute_reward': False, 'alignment_weighting': 0.0, 'oracle_bonus': 0.0,
           'rng': np.random.RandomState(1337), 'device': dev, 'target_sh_order': 8,
           'noise': 0.0, 'fa_map': None, 'state_of_stopped': False}
    env = NoisyTrackingEnvironment(subject, 'testing', dto)
    del sub
    torch.cuda.empty_cache()
    # exactly N seeds on the mask shell, same on every rank (seeded), then the rank's slice
    from tracktolearn_b200.environments.utils import random_seeds_from_mask
    rs = np.random.RandomState(1337)
    seed_mask = subject[2].data if hasattr(subject[2], 'data') else None
    seed_mask = np.asarray(seed_mask.cpu() if hasattr(seed_mask, 'cpu') else seed_mask)
    n_vox = int(seed_mask.astype(bool).sum())
    npv = max(1, -(-a.seeds // n_vox))
    seeds = random_seeds_from_mask(seed_mask, npv, rs)
    rs.shuffle(seeds)
    seeds = seeds[:a.seeds]
    s0, s1 = parallel.shard_bounds(len(seeds), rank, world)
    env.seeds = seeds[s0:s1]
    alg = SACAuto(B.STATE_SIZE, 3, B.HIDDEN, n_actors=a.n_actor, device=dev, precision=a.precision)
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(B.STATE_SIZE, B.HIDDEN, seed=1111, kind='tracking'))
    alg.use_cuda_graph = bool(a.graph)
    tracker = Tracker(alg, a.n_actor, min_length=10.0, max_length=B.MAX_LENGTH_MM)
    stream = torch.cuda.current_stream(dev)
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # one untimed pass first, same seeds: device buffers and pinned staging memory are allocated once per
    # process (a 1 GB cudaHostAlloc inside the timed region would cost more than the tracking)
    for _ in tracker.track_packed(env, copy=False):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    units = steps = d2h = 0
    for batch in tracker.track_packed(env, copy=False):     # pinned staging buffers, consumed at once
        units += env.streamline_steps()
        steps += alg.last_episode_steps
        d2h += batch.data.nbytes + batch.offsets.nbytes
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    # second, untimed pass (tracking is deterministic at prob = 0): keep the streamlines for the
    # property checks and the gather
    parts = [b for b in tracker.track_packed(env, copy=True)]
    local = parts[0]
    for p in parts[1:]:
        local += p
    # size-independent checks at full size
    lens = np.diff(local.offsets)
    ok = {
        'one_streamline_per_seed': bool(len(local) == len(env.seeds)),
        'starts_on_seed': bool(np.array_equal(local.data[local.offsets[:-1]],
                                              env.seeds.astype(np.float32))),
        'lengths_in_range': bool(lens.min() >= 1 and lens.max() <= env.max_nb_steps + 1),
        'all_stopped_with_a_flag': bool((local.data_per_streamline['flags'] != 0).all()),
        'finite_or_nan_mask_stop': bool(np.isfinite(local.data).all()),
    }
    gather_ms = None
    n_total = len(local)
    if world > 1 and not a.no_gather:
        # the final exchange from the device (the env still holds the last pass): every rank's packed
        # streamlines go to rank 0 over NVLink, one pinned D2H there.  The first call also sets up
        # NCCL's point-to-point channels.
        gather_ms = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            merged = parallel.gather_env_streamlines(env, copy=False)
            barrier()
            gather_ms.append(1000.0 * (time.perf_counter() - t0))
        if rank == 0:
            n_total = len(merged)
            ok['gathered_all'] = bool(n_total == a.seeds)
            s0, s1 = parallel.shard_bounds(len(seeds), 0, world)
            ok['gathered_rank0_slice_identical'] = bool(
                np.array_equal(merged.data[:merged.offsets[s1 - s0]], local.data)
                and np.array_equal(merged.data_per_streamline['seeds'], seeds))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    u = torch.tensor([units, steps, float(all(ok.values()))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        umin = u.clone()
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(umin, op=dist.ReduceOp.MIN)
        all_ok = bool(umin[2] > 0)
    else:
        all_ok = all(ok.values())
    if rank == 0:
        print(json.dumps({
            'metric': 'streamline-steps/sec', 'value': float(u[0]) / (float(t[0]) * 1e-3),
            'unit': 'streamline-steps/s', 'n_gpus': world, 'scaling': 'strong',
            'config': {'workload': '0.5mm-iso synthetic %dx%dx%d order-8 fODF, %d seeds sharded across %d GPU(s)'
                                   % (shape + (a.seeds, world)),
                       'n_actor_per_gpu': a.n_actor, 'step_mm': STEP_MM, 'max_nb_steps': int(env.max_nb_steps),
                       'what': 'Tracker.track_packed: seeds H2D, full episodes incl. tail, packed streamlines D2H'},
            'tracking_ms': float(t[0]), 'streamline_steps': float(u[0]), 'env_steps_rank_sum': float(u[1]),
            'streamlines': n_total, 'mean_points': float(lens.mean()),
            'streamlines_per_s': a.seeds / (float(t[0]) * 1e-3),
            'd2h_bytes_rank0': int(d2h), 'tractogram_gather_ms_first_and_warm': gather_ms, 'setup_s_rank0': setup_s,
            'properties_ok': all_ok, 'properties_rank0': ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
