#!/bin/bash
# One 8-GPU session: bench.py (headline tier), configs[2] strong scaling, configs[3] training, configs[4] oracle.
n=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --only-main > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
echo "bench rc=$?"
timeout 400 $TR --master-port 29534 benchmarks/config3_bench.py 2>/dev/null | tail -1 > gpurun_out/config3_n$n.json; echo "config3 rc=$?"
timeout 400 $TR --master-port 29535 benchmarks/sac_train_bench.py 2>/dev/null | tail -1 > gpurun_out/sac_train_n$n.json; echo "sac rc=$?"
timeout 400 $TR --master-port 29536 benchmarks/oracle_bench.py 2>/dev/null | tail -1 > gpurun_out/oracle_n$n.json; echo "oracle rc=$?"
python - <<PY
import json
n = $n
try:
    d = json.load(open('gpurun_out/bench_n%d.json' % n))
    print('bench N=%d value %.1f M  e2e %.1f M (%.1f ms)' % (d['n_gpus'], d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms']))
    s = d['sharded']; print('sharded %.1f M total %.1f ms track %.1f gather %.1f steps %d occ %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['tracking_ms_max_over_ranks'], s['gather_ms'], s['env_steps_max_over_ranks'], s['mean_slot_occupancy'], s['properties_ok']))
except Exception as e: print('bench', e)
for name in ('config3', 'sac_train', 'oracle'):
    try:
        d = json.load(open('gpurun_out/%s_n%d.json' % (name, n)))
        print(name, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items() if k in ('value', 'tracking_ms', 'tractogram_gather_ms_first_and_warm', 'updates_per_s_per_replica', 'ms_per_env_step_plus_update', 'replica_weight_divergence', 'device_resident_streamlines_per_s', 'host_to_host_streamlines_per_s', 'properties_ok', 'n_gpus')})
    except Exception as e: print(name, e)
PY
