#!/bin/bash
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_settle.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'skip':os.environ.get('CONSTV_SETTLE_SKIP'),'args':'$*','us':round(d['ms_per_step']*1e3,3)}))" || tail -5 gpurun_out/err_settle.log; }
run --rigid-water fused
export CONSTV_SETTLE_SKIP=1
run --rigid-water fused
run --rigid-water fused --temperature 0
