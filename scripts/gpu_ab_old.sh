#!/bin/bash
# A/B on one box: the tree of an older commit (unpacked and built under ab_old/) against the current tree.
mkdir -p gpurun_out
run() {  # dir tag
  ( cd $1 && timeout 200 python bench.py --only-main --no-cpu --no-sharded --no-e2e --steps 20 --warmup 5 2>/dev/null ) > gpurun_out/ab_$2.json
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$2.json")); t = d["tiers"]["fp16"]
print("$2 fp16 %.1f M %.1f us" % (d["value"]/1e6, d["ms_per_step"]*1e3), {k: round(v["avg_us"],1) for k,v in t["kernels"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run ab_old old1
run . new1
run ab_old old2
run . new2
