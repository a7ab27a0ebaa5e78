#!/bin/bash
# bench.py under torchrun on N GPUs of the box.   scripts/gpu_multi.sh <tag> <N> [bench args]
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 50 --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench N=$n rc=$?"; tail -c 800 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$tag.json'))
print('N=%d value %.1f M  %.1f us/step  e2e %.1f M (%.1f ms)  launches %d' % (d['n_gpus'], d['value'] / 1e6, d['ms_per_step'] * 1e3, d['e2e']['value'] / 1e6, d['e2e']['ms'], d['gpu_launches']))
print(d['e2e']['what'])
s = d['sharded']
if s: print('sharded %.1f M  total %.1f ms  track %.1f ms  gather %.1f ms  steps %d occupancy %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['tracking_ms_max_over_ranks'], s['gather_ms'], s['env_steps_max_over_ranks'], s['mean_slot_occupancy'], s['properties_ok']))
PY
