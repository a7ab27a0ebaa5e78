#!/usr/bin/env python
"""BASELINE.json configs[3]: sac_auto_train rollout + update loop, n_actor = 4096, alignment reward,
batch 4096, 615-1024-1024-1024 networks, gradient all-reduce across the GPUs of one box.

    python benchmarks/sac_train_bench.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 benchmarks/sac_train_bench.py       # N GPUs (NCCL)

Every rank runs its own environment (n_actor streamlines, its own seeds) and replay shard; after
every environment step one SAC-auto update is done on a batch drawn from the local shard, with ONE
all-reduce per optimiser over the flattened gradients.  Reports environment
streamline-steps/s, updates/s and the replica divergence of the actor weights after the run (must
be 0: replicas stay in lock-step)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--episodes', type=int, default=3)
    ap.add_argument('--n-actor', type=int, default=4096)
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--shape', type=int, nargs=3, default=[64, 64, 64])
    ap.add_argument('--tf32', action='store_true', help='let the learner\'s torch matmuls use TF32 tensor cores '
                    '(the reference and the default run them in fp32)')
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import TrackingEnvironment
    from tracktolearn_b200.environments.utils import random_seeds_from_mask

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    shape = tuple(a.shape)
    sub = synthetic.make_subject(shape, seed=1234, device=dev, with_peaks=True)
    affine = np.eye(4)
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['seed_mask'], affine), MRIDataVolume(sub['peaks'], affine), affine)
    # models/hyperparameters.json: step 0.75 mm, theta 30, max_length 200, n_dirs 100
    dto = {'n_dirs': 100, 'theta': 30.0, 'npv': 2, 'binary_stopping_threshold': 0.1, 'step_size': 0.75,
           'min_length': 20.0, 'max_length': 200.0, 'oracle_checkpoint': None,
           'oracle_stopping_criterion': False, 'scoring_data': None, 'compute_reward': True,
           'alignment_weighting': 1.0, 'oracle_bonus': 0.0, 'rng': np.random.RandomState(1337 + rank),
           'device': dev, 'target_sh_order': 8, 'noise': 0.0, 'fa_map': None}
    env = TrackingEnvironment(subject, 'training', dto)
    rs = np.random.RandomState(1337 + rank)
    env.seeds = random_seeds_from_mask(sub['seed_mask'].cpu().numpy(), 2, rs)
    alg = SACAuto(615, 3, '1024-1024-1024', n_actors=a.n_actor, batch_size=a.batch, device=dev, precision='bf16')
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(615, '1024-1024-1024', seed=1111, kind='tracking'))
    learner = alg.enable_training(replay_size=1000000, batch_size=a.batch, start_timesteps=a.n_actor)
    if a.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
    np.random.seed(rank)
    torch.manual_seed(rank)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up episode (cuBLAS heuristics, Adam state, plans), untimed
    alg._episode(env.nreset(a.n_actor), env)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_before = alg.t
    it_before = alg.total_it
    e0.record(stream)
    lengths, rewards = [], []
    for _ in range(a.episodes):
        r, losses, length, _ = alg._episode(env.nreset(a.n_actor), env)
        lengths.append(length)
        rewards.append(r)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    units = alg.t - t_before
    updates = alg.total_it - it_before
    # replicas in lock-step: the actor's first-layer weights are identical on every rank
    w = learner.actor.layers[0].weight.detach().double()
    chk = torch.stack([w.sum(), w.abs().sum()])
    div = 0.0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    u = torch.tensor([units, updates], dtype=torch.float64, device=dev)
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        div = float((hi - lo).abs().max())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        usum = u.clone()
        dist.all_reduce(usum, op=dist.ReduceOp.SUM)
    else:
        usum = u
    if rank == 0:
        sec = float(t[0]) * 1e-3
        grad_bytes = sum(p.numel() for p in learner.actor.parameters()) * 4 + \
            sum(p.numel() for p in learner.critic.parameters()) * 4 + 4
        print(json.dumps({'learner_matmul': 'tf32' if a.tf32 else 'fp32',
            'metric': 'training streamline-steps/sec (rollout + one SAC update per env step)',
            'value': float(usum[0]) / sec, 'unit': 'streamline-steps/s', 'n_gpus': world, 'scaling': 'weak',
            'updates_per_s_per_replica': float(u[1]) / sec, 'env_steps_rank0': int(sum(lengths)),
            'ms_per_env_step_plus_update': float(t[0]) / max(1, int(sum(lengths))),
            'config': {'workload': 'sac_auto_train rollout+update, %dx%dx%d volume, n_actor=%d per GPU, batch=%d, '
                                   'alignment reward, 615-1024-1024-1024 actor and double critic' %
                                   (shape + (a.n_actor, a.batch)),
                       'episodes': a.episodes, 'allreduce_bytes_per_update': grad_bytes if world > 1 else 0,
                       'collective': 'NCCL all-reduce of flattened gradients, one per optimiser' if world > 1 else None},
            'episode_lengths_rank0': lengths, 'episode_rewards_rank0': rewards,
            'replica_weight_divergence': div}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
