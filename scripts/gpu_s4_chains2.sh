#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_settle.py tests/test_gpu_drude.py -m gpu -x -q 2>&1 | tail -3
: > gpurun_out/sweep_chains2.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_chains2.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" | tee -a gpurun_out/sweep_chains2.jsonl || tail -5 gpurun_out/err_chains2.log; }
for c in 1 2 4; do run --rigid-water fused --chains $c; done
for c in 1 2 4; do run --drude --chains $c; done
run --rigid-water fused --chains 2 --atoms 4000000 --steps 3000
run --drude --chains 2 --atoms 4000000 --steps 3000
run --rigid-water fused --chains 2 --precision mixed
run --drude --chains 2 --precision mixed
