#!/bin/bash
# The ncu captures under profiles/ (B200_PROFILING.md recipe), to be run on the GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/run_ncu.sh r2 fp16'
# 1. the plain command must exit 0 first; 2. launch list of steady-state steps; 3. --set full of the three
# kernels of two steps.  Outputs land in gpurun_out/ (scripts/ncu_summary.py turns the report into the
# summary bench.py reads; copy both into profiles/).
tag=${1:-r2}; prec=${2:-fp16}
CMD="python bench.py --no-cpu --no-e2e --no-sharded --only-main --precision $prec --steps 2 --warmup 3"
K='regex:mlp_pair_kernel|propagate_stop_kernel|build_state_kernel'
# launches before the timed region: (384 burn-in + 3 warm-up) steps x 3 kernels
SKIP=1161
set -x
$CMD > gpurun_out/ncu_plain_$tag.json 2> gpurun_out/ncu_plain_$tag.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip $SKIP -c 30 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
$CMD > gpurun_out/ncu_plain2_$tag.json 2>> gpurun_out/ncu_plain_$tag.err &&
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip $SKIP -c 6 -f -o gpurun_out/step_$tag \
    $CMD > gpurun_out/ncu_f_$tag.log 2>&1
ls -la gpurun_out/step_$tag.ncu-rep gpurun_out/launches_$tag.csv
