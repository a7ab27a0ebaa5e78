#!/bin/bash
# bench.py with the given extra args; prints the per-tier summary.   scripts/gpu_bench_only.sh <tag> [bench args]
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$tag.json'))
print('value %.1f M  %.1f us/step  e2e %.1f M  launches %d' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, (d['e2e'] or {}).get('value', 0) / 1e6, d['gpu_launches']))
for p, t in d['tiers'].items():
    r = t['roofline'] or {}
    print(p, '%.1f M  %.1f us/step  dense %.1f us  %.0f TF  frac %.3f  %s MHz' % (t['value'] / 1e6, t['ms_per_step'] * 1e3, r.get('avg_launch_us', 0), r.get('achieved', 0), r.get('frac', 0), t.get('sm_mhz_timed_region')),
          {k: round(v['avg_us'], 1) for k, v in t['kernels'].items()})
print('clocks', d['clocks'])
if d.get('sharded'):
    s = d['sharded']; print('sharded %.1f M  total %.1f ms  gather %.1f ms  occupancy %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['gather_ms'], s['mean_slot_occupancy'], s['properties_ok']))
PY
