#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_drude.py tests/test_gpu_settle.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_reord.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" | tee -a gpurun_out/sweep_reordered2.jsonl || tail -5 gpurun_out/err_reord.log; }
: > gpurun_out/sweep_reordered2.jsonl
run --drude
run --drude --reordered
run --drude --reordered --variant 1
run --rigid-water fused --reordered
run --rigid-water fused --reordered --variant 1
run --reordered
run --drude --reordered --atoms 4000000 --steps 3000
run --rigid-water fused --reordered --atoms 4000000 --steps 3000
