#!/bin/bash
# full GPU suite, smoke, default bench (both arms) -- the round-end sequence of the driver
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s4_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_s4_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/s4f_bench_reference.json 2>gpurun_out/s4f_err.log; tail -c 600 gpurun_out/s4f_bench_reference.json
python bench.py 2>>gpurun_out/s4f_err.log | tail -1 > gpurun_out/s4f_bench_default.json
python -c "
import json; d=json.load(open('gpurun_out/s4f_bench_default.json'))
print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
