#!/usr/bin/env python
"""bench.py -- streamline-steps/sec of the batched tracking step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): whole-brain synthetic 145x174x145 1.25 mm order-8
descoteaux07 fODF, npv 20 on the mask shell (~1M seeds per GPU), n_actor 50 000,
NoisyTrackingEnvironment with noise 0 (what ttl_track.py always runs), SAC actor
615-1024-1024-1024-6 with a synthetic "tracking-like" checkpoint.  One bench STEP = one pass of
the hot path over the batch of n_actor alive streamlines: actor forward (state pack, three
tcgen05 dense layers with the head fused into the last, head finish) + env step (propagate/stop
with ordered-compaction bookkeeping and slot refill, state gather) -- 6 kernel launches, no host
involvement.

The JSON line follows the driver contract; see DESIGN.md section "Measurement" for how every
field is produced.  `--impl reference` times the CPU restatement of the reference path
(oracle/ttl_oracle.py + a torch-CPU actor with all host threads) on a bounded sample of the
same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (145, 174, 145)
VOXEL_MM = 1.25
NPV = 20
N_ACTOR = 50000
TRAINED_VOXEL = 0.9987237            # models/hyperparameters.json:49
TRAINED_STEP = 0.75
STEP_MM = VOXEL_MM / TRAINED_VOXEL * TRAINED_STEP   # runners/ttl_track.py:126-141
MAX_LENGTH_MM = 300.0                # ttl_track.py --max_length default
THETA = 30.0
HIDDEN = '1024-1024-1024'
STATE_SIZE = 615
ACTOR_FLOP_PER_ROW = 2 * (615 * 1024 + 2 * 1024 * 1024 + 1024 * 6)      # 5 466 112
DENSE_FLOP_PER_ROW = 2 * (615 * 1024 + 2 * 1024 * 1024)                 # the three tcgen05 layers
STEP_BYTES_PER_ROW = 8900            # gather 4680 + state 2460 + dirs 1200 + mask 512 + 48
STATE_KERNEL_BYTES_PER_ROW = 4680 + 2460 + 1200 + 12
WORKLOAD = 'whole-brain synthetic 145x174x145 1.25mm order-8 fODF, npv=20, n_actor=50000'
CPU_SAMPLE_ROWS = 4096
BURN_IN = 128
CONFIG = 2


def set_config(n):
    """BASELINE.json configs[1] (default, the configuration the metric is quoted on) or configs[2]
    (0.5 mm-iso 290^3 volume, 798 steps max, ~1M seeds per GPU) for the steady-state legs."""
    global SHAPE, VOXEL_MM, NPV, STEP_MM, WORKLOAD, CONFIG
    CONFIG = n
    if n == 3:
        SHAPE, VOXEL_MM, NPV = (290, 290, 290), 0.5, 5
        STEP_MM = VOXEL_MM / TRAINED_VOXEL * TRAINED_STEP
        WORKLOAD = '0.5mm-iso synthetic 290x290x290 order-8 fODF, npv=5 (~1M seeds), n_actor=50000'


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the step's kernels from
    the committed `ncu --set full` capture of this same command (profiles/README.md); None when the
    summary file is absent."""
    path = os.path.join(ROOT, 'profiles', 'r1_step_v9_ncu_summary.json')
    if not os.path.exists(path):
        return {}

    def mbytes(txt):
        v, unit = txt.split()[:2]
        return float(v) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    per = {}
    with open(path) as f:
        for e in json.load(f):
            name = e['kernel'].split('<')[0]
            per.setdefault(name, []).append(mbytes(e['dram__bytes_read.sum']) + mbytes(e['dram__bytes_write.sum']))
    return {k: sum(v) / len(v) for k, v in per.items()}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '10'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def wait_first(self, timeout_s):
        """Block until nvidia-smi has produced its first sample (it starts slowly)."""
        t0 = time.monotonic()
        while self.proc is not None and not self.rows and time.monotonic() - t0 < timeout_s:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def mark_end(self):
        self.t_end = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0, t1 = getattr(self, 't_begin', 0.0), getattr(self, 't_end', float('inf'))
        # a sample is read ~one period after it was taken: accept rows up to 40 ms past the end mark;
        # a timed region shorter than the sampling period falls back to the nearest samples around it
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.04]
        window = 'timed region'
        if len(rows) < 2:
            rows = [r for (t, r) in self.rows if t0 - 0.1 <= t <= t1 + 0.1]
            window = 'timed region +-100 ms (region shorter than the sampling period)'
        sm, smax, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                 f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference path restated (oracle) with a torch-CPU actor on all host threads
# ------------------------------------------------------------------------------------------
def cpu_reference_run(sub_np, seeds, actor_sd, rows, steps, warmup):
    import torch
    from oracle import ttl_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    W = [(actor_sd['layers.%d.weight' % (2 * i)].float(), actor_sd['layers.%d.bias' % (2 * i)].float())
         for i in range(4)]

    def torch_actor(state):      # offpolicy.py:94-140 at probabilistic = 0
        with torch.no_grad():
            h = torch.from_numpy(state)
            for i, (w, b) in enumerate(W):
                h = torch.addmm(b, h, w.t())
                if i < 3:
                    h = torch.relu(h)
            return torch.tanh(h[:, :3]).numpy()

    env = O.OracleEnv(sub_np['sh'], sub_np['mask'], seeds, VOXEL_MM, STEP_MM, theta=THETA,
                      max_length_mm=MAX_LENGTH_MM, noisy=True)
    state = env.reset(0, rows)
    total, t_total, it, start_pos = 0, 0.0, 0, rows
    per_step = []
    while it < warmup + steps:
        if len(env.continue_idx) == 0:
            end = min(start_pos + rows, len(seeds))
            if end <= start_pos:
                start_pos, end = 0, rows
            state = env.reset(start_pos, end)
            start_pos = end
        t0 = time.perf_counter()
        n = len(env.continue_idx)
        action = torch_actor(state)
        env.step(action)
        state, _ = env.harvest()
        dt = time.perf_counter() - t0
        if it >= warmup:
            total += n
            t_total += dt
            per_step.append(dt)
        it += 1
    return total / t_total, threads, 1000.0 * t_total / max(1, len(per_step))


def build_subject_numpy():
    from tracktolearn_b200 import synthetic
    sub = synthetic.make_subject(SHAPE, seed=1234, with_peaks=False)
    return {k: (v.numpy() if v is not None else None) for k, v in sub.items()}


def draw_seeds(seed_mask, rank):
    from tracktolearn_b200.environments.utils import random_seeds_from_mask
    rs = np.random.RandomState(1337 + rank)
    seeds = random_seeds_from_mask(seed_mask, NPV, rs)
    rs.shuffle(seeds)
    return seeds


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import torch
    from tracktolearn_b200 import synthetic
    sub = build_subject_numpy()
    seeds = draw_seeds(sub['seed_mask'], 0)
    sd = synthetic.actor_state_dict(STATE_SIZE, HIDDEN, seed=1111, kind='tracking')
    value, threads, ms = cpu_reference_run(sub, seeds, sd, CPU_SAMPLE_ROWS, args.steps, args.warmup)
    sample = ('%d act->step->harvest iterations over a %d-streamline batch of the same volume/seeds '
              '(numpy/scipy env restatement + torch-CPU fp32 actor)' % (args.steps, CPU_SAMPLE_ROWS))
    line = {
        'impl': 'reference', 'metric': 'streamline-steps/sec', 'value': value, 'unit': 'streamline-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample_rows': CPU_SAMPLE_ROWS, 'impl': 'oracle port on host cores'},
        'cpu_baseline': {'value': value, 'unit': 'streamline-steps/s', 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'streamline-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on stdout from C): point fd 1 at stderr for the run and keep the real stdout for the line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main_gpu(args):
    quiet_stdout()
    import torch
    import torch.distributed as dist
    from tracktolearn_b200 import _lib, synthetic
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    from tracktolearn_b200.datasets.utils import MRIDataVolume
    from tracktolearn_b200.environments import NoisyTrackingEnvironment
    from tracktolearn_b200.tracking.tracker import Tracker

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    # ---- workload ---------------------------------------------------------------------
    sub = synthetic.make_subject(SHAPE, seed=1234, device=dev, with_peaks=False)
    affine = np.diag([VOXEL_MM] * 3 + [1.0])
    subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine),
               MRIDataVolume(sub['seed_mask'], affine), None, affine)
    dto = {'n_dirs': 100, 'theta': THETA, 'npv': 1, 'binary_stopping_threshold': 0.1,
           'step_size': STEP_MM, 'min_length': 10.0, 'max_length': MAX_LENGTH_MM,
           'oracle_checkpoint': None, 'oracle_stopping_criterion': False, 'scoring_data': None,
           'compute_reward': False, 'alignment_weighting': 0.0, 'oracle_bonus': 0.0,
           'rng': np.random.RandomState(1337), 'device': dev, 'target_sh_order': 8,
           'noise': 0.0, 'fa_map': None, 'state_of_stopped': False}
    env = NoisyTrackingEnvironment(subject, 'testing', dto)
    seed_mask_np = sub['seed_mask'].cpu().numpy()
    env.seeds = draw_seeds(seed_mask_np, rank)         # npv=20 per rank: weak scaling
    n_seeds = len(env.seeds)
    actor_sd = synthetic.actor_state_dict(STATE_SIZE, HIDDEN, seed=1111, kind='tracking')
    alg = SACAuto(STATE_SIZE, 3, HIDDEN, n_actors=N_ACTOR, device=dev, precision='bf16')
    alg.agent.actor.load_state_dict(actor_sd)
    actor = alg.agent.actor
    stream = torch.cuda.current_stream(dev)

    from tracktolearn_b200.algorithms.rl import StepRunner
    runner_box = {}

    def one_step(action_buf=None):
        # the same enqueue-only iteration Tracker / validation_episode run (CUDA-graph replay
        # after two warm iterations)
        runner_box['r'].step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput: inputs already in HBM ----------------------------
    env.reset_streaming(0, n_seeds, N_ACTOR, fp32_state=not args.bf16_state_only)
    action_buf = None
    runner_box['r'] = StepRunner(env, actor, 0.0, use_graph=args.graph)
    # burn-in (untimed, part of preparing the workload): with slot refill the alive set needs about
    # two mean lifetimes to reach its steady-state mix of streamline ages and positions; right after
    # reset every streamline still sits on the seed shell and the gather enjoys unrepresentative L2
    # locality
    # the clock sampler starts before the untimed steps: nvidia-smi needs 0.1-0.3 s before its first
    # sample, longer than a short timed region; only the samples taken between the two marks count
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(BURN_IN):
        one_step(action_buf)
    for _ in range(args.warmup):
        one_step(action_buf)
    env.n_alive()
    steps_before = env.streamline_steps()
    launches_before = lib.ttl_launch_count()
    replays_before = runner_box['r'].replays
    barrier()
    sampler.wait_first(2.0)
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        one_step(action_buf)
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    # kernels launched one by one + kernels replayed from the captured step graphs
    gpu_launches = int(lib.ttl_launch_count() - launches_before) + \
        (runner_box['r'].replays - replays_before) * runner_box['r'].kernels_per_step
    env.n_alive()
    units = env.streamline_steps() - steps_before
    alive_end = int(env._batch.ctrl_host[env._cur])

    # ---- per-kernel device times (CUDA events on the launching stream) -------------------
    prof_steps = min(20, args.steps)
    lib.ttl_prof_enable(1)
    plain = StepRunner(env, actor, 0.0, use_graph=False)
    for _ in range(prof_steps):
        plain.step()
    torch.cuda.synchronize(dev)
    prof = _lib.prof_report()
    lib.ttl_prof_enable(0)
    env.n_alive()
    rows_prof = alive_end   # alive count is pinned at n_actor while seeds remain

    # ---- end to end through the public API, host buffers ------------------------------------
    e2e_seeds = min(n_seeds, 16 * N_ACTOR) if not args.no_e2e else N_ACTOR // 50
    env_seeds_all = env.seeds
    env.seeds = env_seeds_all[:e2e_seeds]
    tracker = Tracker(alg, N_ACTOR, min_length=10.0, max_length=MAX_LENGTH_MM)
    # one untimed pass first: device buffers and pinned staging memory are allocated once per
    # process (W >= 3 warm-up rule applies to the e2e leg too), then the same call is timed
    for batch in tracker.track_packed(env, copy=False):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    d2h = 0
    e2e_units = 0
    e2e_steps = 0
    n_streamlines = 0
    for batch in tracker.track_packed(env, copy=False):   # seeds H2D, episodes, packed streamlines D2H
        d2h += batch.data.nbytes + batch.offsets.nbytes + batch.data_per_streamline['flags'].nbytes
        e2e_units += env.streamline_steps()
        e2e_steps += alg.last_episode_steps
        n_streamlines += len(batch)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = e2e_seeds * 3 * 8
    env.seeds = env_seeds_all

    # ---- reduce over ranks: max time, summed units -------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([units, e2e_units], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    units_all, e2e_units_all = float(u[0]), float(u[1])
    value = units_all / (elapsed_ms * 1e-3)
    e2e_value = e2e_units_all / (e2e_ms * 1e-3)

    pk = peaks()

    def avg_ms(name):
        n, ms = prof.get(name, (0, 0.0))
        return ms / n if n else None

    # the three tcgen05 layers of one step (the last one carries the fused head)
    dn = [prof.get(k, (0, 0.0)) for k in ('dense_bf16_kernel', 'dense_bf16_head_kernel')]
    dense_launches = sum(n for n, _ in dn)
    dense_total_ms = sum(ms for _, ms in dn)
    traffic = ncu_traffic()
    traffic_src = 'profiles/r1_step_v9_ncu_summary.json (ncu --set full of this command, bytes per launch)'
    roofline = None
    if dense_launches:
        steps_prof = dense_launches / 3.0
        achieved = DENSE_FLOP_PER_ROW * rows_prof * steps_prof / (dense_total_ms * 1e-3) / 1e12
        roofline = {'kernel': 'dense_bf16_kernel (tcgen05 actor layers, 3 launches/step, last one with fused head)',
                    'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_tflops_sustained'],
                    'unit': 'TFLOP/s', 'frac': achieved / pk['bf16_tflops_sustained'],
                    'traffic': traffic.get('dense_bf16_2cta_kernel'), 'traffic_source': traffic_src,
                    'avg_launch_us': 1000.0 * dense_total_ms / dense_launches,
                    'flop_per_launch': DENSE_FLOP_PER_ROW * rows_prof / 3.0,
                    'peak_source': pk['source'] + ', sustained bf16 figure (kernel timed inside a long step)'}
    kernels = {}
    for name, (n, ms) in sorted(prof.items()):
        kernels[name] = {'launches': n, 'avg_us': 1000.0 * ms / n}
    # algorithmic bytes per row; without the fp32 API tensor the state write is the
    # 1280-byte bf16 operand instead of the 2460-byte fp32 row
    state_write = 1280 if args.bf16_state_only else 2460 + 1280
    # previous directions: 100 fp32 points re-read (API mode), or the previous row's 600-byte bf16
    # direction block shifted by one direction (device mode)
    dirs_read = 600 if args.bf16_state_only else 1200
    state_bytes = 4680 + state_write + dirs_read + 12
    step_bytes = 4680 + state_write + dirs_read + 512 + 48
    state_ms = avg_ms('build_state_kernel')
    step_ms = sum(avg_ms(k) or 0.0 for k in ('propagate_stop_kernel', 'build_state_kernel'))
    roofline_step = None
    if state_ms and step_ms:
        a_state = state_bytes * rows_prof / (state_ms * 1e-3) / 1e9
        a_step = step_bytes * rows_prof / (step_ms * 1e-3) / 1e9
        roofline_step = {
            'build_state_kernel': {'bound': 'hbm', 'achieved': a_state, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                   'frac': a_state / pk['hbm_gbs'], 'traffic': traffic.get('build_state_kernel'),
                                   'algorithmic_bytes_per_launch': state_bytes * rows_prof,
                                   'bytes_per_row': state_bytes},
            'env_step (propagate_stop+build_state)': {
                'bound': 'hbm', 'achieved': a_step, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                'frac': a_step / pk['hbm_gbs'],
                'traffic': (traffic.get('build_state_kernel', 0) + traffic.get('propagate_stop_kernel', 0)) or None,
                'algorithmic_bytes_per_launch': step_bytes * rows_prof, 'bytes_per_row': step_bytes},
            'peak_source': pk['source']}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sub_np = {'sh': sub['sh'].cpu().numpy(), 'mask': sub['mask'].cpu().numpy()}
        v, threads, ms = cpu_reference_run(sub_np, env.seeds, actor_sd, CPU_SAMPLE_ROWS, 24, 3)
        cpu_baseline = {'value': v, 'unit': 'streamline-steps/s', 'cores': threads, 'kind': 'port',
                        'sample': '24 act->step->harvest iterations (after 3 warm-up) over a %d-streamline '
                                  'batch of the same volume and seeds; numpy/scipy env restatement + torch-CPU '
                                  'fp32 actor; scipy map_coordinates is single-threaded' % CPU_SAMPLE_ROWS,
                        'ms_per_step': ms}

    if rank == 0:
        line = {
            'metric': 'streamline-steps/sec', 'value': value, 'unit': 'streamline-steps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'volume': list(SHAPE), 'voxel_mm': VOXEL_MM, 'npv': NPV,
                       'n_actor': N_ACTOR, 'seeds_per_gpu': n_seeds, 'step_mm': STEP_MM,
                       'max_nb_steps': int(env.max_nb_steps), 'actor': '615-' + HIDDEN + '-6 (synthetic, tracking-like)',
                       'env': 'NoisyTrackingEnvironment noise=0 (float64 directions)',
                       'streaming_refill': True,
                       'slot_order': 'seeds enter the slots in voxel raster order (rows / output keep the shuffled order)'
                       if os.environ.get('TTL_LOCALITY', '1') != '0' else 'row (shuffled) order',
                       'launch': 'programmatic dependent launch' if os.environ.get('TTL_PDL', '1') != '0' else 'plain', 'alive_at_end': alive_end, 'burn_in_steps': BURN_IN,
                       'state_rows': ('bf16 actor operand only; the fp32 API tensor is not materialised in the '
                                      'device loop (SURVEY 7 step 7)' if args.bf16_state_only
                                      else 'fp32 API tensor + bf16 actor operand'),
                       'l2': 'inputs larger than L2: %d MB SH volume + 2x123 MB state rows + 210 MB activations'
                             % (SHAPE[0] * SHAPE[1] * SHAPE[2] * 48 * 4 // 1000000),
                       'parallelism': 'seeds sharded, volume replicated, no data-path collective'},
            'e2e': {'value': e2e_value, 'unit': 'streamline-steps/s',
                    'h2d_bytes_per_step': h2d / max(1, e2e_steps), 'd2h_bytes_per_step': d2h / max(1, e2e_steps),
                    'what': 'Tracker.track_packed over %d seeds per GPU: pinned-host seeds H2D, full episodes '
                            'incl. tail, packed streamlines+flags D2H; %d env steps, %d streamlines'
                            % (e2e_seeds, e2e_steps, n_streamlines)},
            'gpu_launches': gpu_launches,
            'clocks': clocks,
            'roofline': roofline,
            'roofline_step_kernels': roofline_step,
            'kernels': kernels,
            'cpu_baseline': cpu_baseline,
            'flop_per_streamline_step': ACTOR_FLOP_PER_ROW,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--fp32-state', dest='bf16_state_only', action='store_false',
                    help='also materialise the fp32 state rows every step (the reference API tensor)')
    ap.add_argument('--graph', action='store_true', help='replay the kernels of a step from a CUDA graph')
    ap.add_argument('--no-e2e', action='store_true', help='skip the end-to-end leg (profiling runs)')
    ap.add_argument('--config', type=int, default=2, choices=[2, 3],
                    help='2: BASELINE.json configs[1] (default); 3: configs[2], the 290^3 0.5 mm volume')
    a = ap.parse_args()
    set_config(a.config)
    if a.warmup < 3:
        a.warmup = 3
    sys.exit(main_reference(a) if a.impl == 'reference' else main_gpu(a))
