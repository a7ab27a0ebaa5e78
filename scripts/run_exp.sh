run() { tag=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 400 --warmup 20 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/exp_$tag.json'));print('$tag', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/exp_$tag.err; }
run dedup_occ3 TTL_STATE_OPTIONS=32
run dedup_occ4 TTL_STATE_OPTIONS=64
run corner56 TTL_STATE_OPTIONS=16
