#!/bin/bash
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_ens.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" | tee -a gpurun_out/sweep_ens_water.jsonl || tail -5 gpurun_out/err_ens.log; }
: > gpurun_out/sweep_ens_water.jsonl
for c in 1 2 4 8; do run --ensemble 8 --atoms 100000 --rigid-water fused --chains $c; done
run --ensemble 8 --atoms 100000 --chains 4
