#!/bin/bash
# session 4: rigid-molecule (SETTLE) step: GPU parity, plugin test, fused vs split timing
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_settle.py tests/test_plugin_cpp.py -m gpu -x -q > gpurun_out/pytest_gpu_settle.log 2>&1
tail -25 gpurun_out/pytest_gpu_settle.log
: > gpurun_out/sweep_settle.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_settle.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep_settle.jsonl || tail -5 gpurun_out/err_settle.log; }
run --rigid-water fused
run --rigid-water split
run --chains 1
run --rigid-water fused --atoms 4000000 --steps 3000
run --rigid-water split --atoms 4000000 --steps 3000
run --chains 1 --atoms 4000000 --steps 3000
run --rigid-water fused --precision mixed
run --rigid-water split --precision mixed
