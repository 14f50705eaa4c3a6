#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep_precision.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_prec.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4),'bytes_per_step':d['roofline']['algorithmic_bytes_per_step']}))" \
    | tee -a gpurun_out/sweep_precision.jsonl || tail -5 gpurun_out/err_prec.log; }
for p in single mixed double; do run --precision $p --chains 1; run --precision $p --chains 2; done
run --atoms 100000 --chains 1
run --atoms 100000 --chains 2
run --atoms 15000 --chains 1
run --atoms 15000 --chains 1 --keep-in-l2 --replicas 1
run --ensemble 8 --atoms 100000 --chains 4
run --ensemble 64 --atoms 100000 --chains 4 --steps 3000
