#!/bin/bash
n=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --only-main > gpurun_out/bench_n${n}b.json 2> gpurun_out/bench_n${n}b.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n${n}b.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_n${n}b.json'))
print('bench N=%d value %.1f M  e2e %.1f M (%.1f ms)' % (d['n_gpus'], d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms']))
s = d['sharded']; print('sharded %.1f M total %.1f ms track %.1f gather %.1f steps %d occ %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['tracking_ms_max_over_ranks'], s['gather_ms'], s['env_steps_max_over_ranks'], s['mean_slot_occupancy'], s['properties_ok']))
PY
timeout 600 python benchmarks/multi_gpu_cli_check.py 2 2>&1 | tail -2
