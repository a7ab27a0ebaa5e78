set -x
python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_plain_v6.json 2> gpurun_out/ncu_plain_v6.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_bf16|propagate_stop|build_state|head_finish" --launch-skip 500 -c 40 --csv --log-file gpurun_out/launches_v6.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_l6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dense_bf16|propagate_stop|build_state" --launch-skip 600 -c 5 -f -o gpurun_out/step_v6 python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_f6.log 2>&1
ls -la gpurun_out/step_v6.ncu-rep; tail -2 gpurun_out/ncu_f6.log
