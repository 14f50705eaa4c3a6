#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
: > gpurun_out/sweep6.jsonl
run() { python bench.py --steps 6000 --warmup 300 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep6.jsonl || tail -3 gpurun_out/err.log; }
for ilp in 1 2 4; do run --pairs-per-thread $ilp; done
run --pairs-per-thread 1 --keep-in-l2
run --pairs-per-thread 1 --no-cache-image-charge
run --pairs-per-thread 1 --no-graph
run --pairs-per-thread 1 --ctas-per-sm 8
run --pairs-per-thread 1 --ctas-per-sm 16
for ilp in 1 2; do run --pairs-per-thread $ilp --atoms 4000000 --steps 3000; done
for ilp in 1 2 4; do run --pairs-per-thread $ilp --atoms 16000000 --steps 1000; done
run --pairs-per-thread 1 --atoms 16000000 --steps 1000 --keep-in-l2
run --pairs-per-thread 1 --atoms 2000000
run --pairs-per-thread 1 --atoms 200000
run --pairs-per-thread 1 --keep-in-l2 --replicas 1
