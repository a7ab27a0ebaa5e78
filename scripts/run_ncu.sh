ncu --set full --clock-control none --import-source on -k regex:"build_state" --launch-skip 130 -c 1 -f -o gpurun_out/state_dedup python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_sd.log 2>&1
TTL_STATE_OPTIONS=32 ncu --set full --clock-control none --import-source on -k regex:"build_state" --launch-skip 130 -c 1 -f -o gpurun_out/state_dedup_occ3 python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_sd3.log 2>&1
ls -la gpurun_out/state_dedup*.ncu-rep
