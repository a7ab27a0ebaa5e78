run() { tag=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 200 --warmup 20 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/exp_$tag.json'));print('$tag', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/exp_$tag.err; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run loc1 TTL_LOCALITY=1
run loc0 TTL_LOCALITY=0
