#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_settle.py -m gpu -x -q > gpurun_out/pytest_gpu_settle.log 2>&1
tail -5 gpurun_out/pytest_gpu_settle.log
: > gpurun_out/sweep_settle2.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_settle.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep_settle2.jsonl || tail -5 gpurun_out/err_settle.log; }
run --rigid-water fused
run --rigid-water fused --keep-in-l2
run --rigid-water fused --temperature 0
run --rigid-water split
run --rigid-water fused --atoms 4000000 --steps 3000
run --rigid-water fused --precision mixed
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --rigid-water fused"
$CMD > gpurun_out/settle_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_settle_step_kernel -s 14 -c 2 -f -o gpurun_out/s4_prof_settle $CMD > gpurun_out/settle_ncu_full.log 2>&1
tail -2 gpurun_out/settle_ncu_full.log
