#!/bin/bash
# End-of-round check on the GPU box: the GPU test suite, smoke, the default bench line, the reference arm.
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$tag.json'))
print('value %.1f M  %.1f us/step  e2e %.1f M  launches %d' % (d['value'] / 1e6, d['ms_per_step'] * 1e3, d['e2e']['value'] / 1e6, d['gpu_launches']))
for p, t in d['tiers'].items():
    r = t['roofline'] or {}
    print(p, '%.1f M  %.1f us/step  dense %.1f us  %.0f TF  frac %.3f  %s MHz' % (t['value'] / 1e6, t['ms_per_step'] * 1e3, r.get('avg_launch_us', 0), r.get('achieved', 0), r.get('frac', 0), t.get('sm_mhz_timed_region')),
          {k: round(v['avg_us'], 1) for k, v in t['kernels'].items()})
print('clocks', d['clocks'])
s = d['sharded']; print('sharded %.1f M  total %.1f ms  gather %.1f ms  occupancy %.2f ok=%s' % (s['value'] / 1e6, s['total_ms'], s['gather_ms'], s['mean_slot_occupancy'], s['properties_ok']))
c = d['cpu_baseline']; print('cpu', round(c['value']), c['cores'], c['threads'], {k: round(v, 2) for k, v in c['phase_seconds'].items()})
print('step roofline', {k: round(v['frac'], 3) for k, v in d['roofline_step_kernels'].items() if isinstance(v, dict)})
PY
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${tag}_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_${tag}_reference.json
