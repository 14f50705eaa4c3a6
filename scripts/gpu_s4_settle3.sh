#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep_settle3.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_settle.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'minb':os.environ.get('CONSTV_SETTLE_MINB'),'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep_settle3.jsonl || tail -5 gpurun_out/err_settle.log; }
for m in 0 6 8 12; do export CONSTV_SETTLE_MINB=$m; run --rigid-water fused; done
export CONSTV_SETTLE_MINB=8; run --rigid-water fused --atoms 4000000 --steps 3000
