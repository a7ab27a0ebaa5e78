run() { tag=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 400 --warmup 20 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/exp_$tag.json'));print('$tag', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), {k:round(v['avg_us'],1) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/exp_$tag.err; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run scan A=1
run scan2 A=1
