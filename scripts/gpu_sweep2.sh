#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/gauss_error.py 2>&1 | tail -6 | tee gpurun_out/gauss_error.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
: > gpurun_out/sweep2.jsonl
run() { python bench.py --steps 6000 --warmup 300 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep2.jsonl || tail -3 gpurun_out/err.log; }
for c in 4 6 8; do run --variant 1 --pairs-per-thread 1 --ctas-per-sm $c; done
for c in 3 4; do run --variant 1 --pairs-per-thread 2 --ctas-per-sm $c; done
for shape in "128 4" "256 2" "256 3" "256 4" "512 2" "512 3"; do set -- $shape
  run --variant 2 --tma-block $1 --tma-stages $2
done
run --variant 2 --tma-block 256 --tma-stages 3 --atoms 4000000
run --variant 1 --pairs-per-thread 1 --ctas-per-sm 8 --atoms 4000000
run --variant 1 --pairs-per-thread 1 --ctas-per-sm 8 --atoms 16000000 --steps 1000
run --variant 1 --pairs-per-thread 1 --ctas-per-sm 8 --no-graph
