#!/bin/bash
# The bench exactly as the round-end driver launches it (BENCH_r01.json: "bench.py --gpus 1 --steps 20 --warmup 5").
tag=${1:-drv}
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo rc=$?
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("value %.1f M  %.1f us  e2e %.1f M (%.1f ms)" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["ms"]))
for p, t in d["tiers"].items():
    print(p, "%.1f M %.1f us (sustained %s)" % (t["value"]/1e6, t["ms_per_step"]*1e3, t["sustained"] and round(t["sustained"]["value"]/1e6,1)), t["sm_mhz_timed_region"], round(t["roofline"]["frac"],3), round(t["roofline"]["avg_launch_us"],1), {k: round(v["avg_us"],1) for k,v in t["kernels"].items()})
s = d["sharded"]; print("sharded %.1f M total %.1f ms gather %.1f" % (s["value"]/1e6, s["total_ms"], s["gather_ms"]))
print(d["run"]["tf32_peak_tflops_measured"], d["clocks"], round(d["cpu_baseline"]["value"]))
PY
