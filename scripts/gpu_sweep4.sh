#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
: > gpurun_out/sweep4.jsonl
run() { python bench.py --steps 3000 --warmup 300 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep4.jsonl || tail -3 gpurun_out/err.log; }
run --atoms 16000000 --steps 1000
run --atoms 16000000 --steps 1000 --temperature 0
run --atoms 16000000 --steps 1000 --temperature 0 --variant 2 --tma-block 256 --tma-stages 2
run --atoms 16000000 --steps 1000 --no-cache-image-charge
run
run --temperature 0
run --temperature 0 --variant 2 --tma-block 256 --tma-stages 2
