run() { tag=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 200 --warmup 20 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/exp_$tag.json'));print('$tag', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})"; }
run base A=1
run pf1 TTL_STATE_PREFETCH=1
run pf2 TTL_STATE_PREFETCH=2
run sorted TTL_BENCH_SORTED_SEEDS=1
run sorted_pf1 TTL_BENCH_SORTED_SEEDS=1 TTL_STATE_PREFETCH=1
