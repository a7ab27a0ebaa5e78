#!/bin/bash
# Run-to-run variance of the single-leg bench on one box: same command three times, then bf16.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,power.limit,clocks.sm,clocks.mem,pstate --format=csv
for i in 1 2 3; do
  timeout 200 python bench.py --only-main --no-cpu --no-sharded --no-e2e --steps 20 --warmup 5 > gpurun_out/rep_$i.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/rep_$i.json")); t = d["tiers"]["fp16"]
print("run $i fp16 %.1f M %.1f us" % (d["value"]/1e6, d["ms_per_step"]*1e3), {k: round(v["avg_us"],1) for k,v in t["kernels"].items()}, d["clocks"])
PY
done
timeout 200 python bench.py --only-main --precision bf16 --no-cpu --no-sharded --no-e2e --steps 20 --warmup 5 > gpurun_out/rep_bf16.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/rep_bf16.json")); t = d["tiers"]["bf16"]
print("bf16 %.1f M %.1f us" % (d["value"]/1e6, d["ms_per_step"]*1e3), {k: round(v["avg_us"],1) for k,v in t["kernels"].items()}, d["clocks"])
PY
nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv
