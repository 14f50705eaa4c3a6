#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_plugin_cpp.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_reord.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" || tail -5 gpurun_out/err_reord.log; }
run --reordered
run --reordered --ctas-per-sm 8
run --reordered --atoms 4000000 --steps 3000
run --reordered --precision mixed
