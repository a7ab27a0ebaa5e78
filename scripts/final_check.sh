#!/bin/bash
# End-of-round check on the GPU box: the GPU test suite, the default bench line, the reference arm.
tag=${1:-final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 300 gpurun_out/bench_$tag.err
python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print(round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), {k:round(v['avg_us'],1) for k,v in d['kernels'].items()}, d['clocks'], d['gpu_launches'], round(d['cpu_baseline']['value']))"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_${tag}_reference.json 2>/dev/null; cut -c1-160 gpurun_out/bench_${tag}_reference.json
python __graft_entry__.py smoke 2>&1 | tail -1
