#!/bin/bash
# One GPU-box session: the GPU test suite, smoke, a short bench line.  Every stage has its own timeout so a
# protocol bug cannot eat the box; logs go to gpurun_out/ (merged back by gpurun).
#   scripts/gpu_check.sh <tag> [pytest args...]
tag=${1:-check}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 "$@" > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$tag.json'))
    print('value %.1f M  %.1f us/step  e2e %.1f M  launches %d' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, (d['e2e'] or {}).get('value', 0) / 1e6, d['gpu_launches']))
    for p, t in d['tiers'].items():
        r = t['roofline'] or {}
        print(p, '%.1f M  %.1f us/step  dense %.1f us  %.0f TF  frac %.3f  %s MHz' % (t['value'] / 1e6, t['ms_per_step'] * 1e3, r.get('avg_launch_us', 0), r.get('achieved', 0), r.get('frac', 0), t.get('sm_mhz_timed_region')),
              {k: round(v['avg_us'], 1) for k, v in t['kernels'].items()})
    print('clocks', d['clocks'])
    if d.get('sharded'):
        s = d['sharded']; print('sharded %.1f M  total %.1f ms  gather %.1f ms  occupancy %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['gather_ms'], s['mean_slot_occupancy'], s['properties_ok']))
except Exception as e:
    print('no bench line:', e)
PY
