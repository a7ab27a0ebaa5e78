#!/bin/bash
# The ncu captures under profiles/ (B200_PROFILING.md recipe), to be run on the GPU box:
#   gpurun --timeout 900 -- 'bash scripts/run_ncu.sh v7'
# 1. the plain command must exit 0 first; 2. launch list of a steady-state step; 3. --set full of the
# kernels of one step.  Outputs land in gpurun_out/ (copy the summaries into profiles/).
tag=${1:-vN}
set -x
python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_plain_$tag.json 2> gpurun_out/ncu_plain_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"dense_bf16|propagate_stop|build_state|head_finish" --launch-skip 500 -c 40 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"dense_bf16|propagate_stop|build_state" --launch-skip 600 -c 5 -f -o gpurun_out/step_$tag \
    python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_f_$tag.log 2>&1
ls -la gpurun_out/step_$tag.ncu-rep
