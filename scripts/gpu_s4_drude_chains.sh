#!/bin/bash
# session 4: GPU parity of the Drude path and of chains, then the chains sweep on the 1M slab
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s4.log 2>&1
tail -15 gpurun_out/pytest_gpu_s4.log
: > gpurun_out/sweep_chains.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e --skip-ref-gpu "$@" 2>gpurun_out/err_chains.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4)}))" \
    | tee -a gpurun_out/sweep_chains.jsonl || tail -3 gpurun_out/err_chains.log; }
run
for c in 2 3 4 6 8 16; do run --chains $c; done
run --chains 4 --no-graph
run --atoms 200000
run --atoms 200000 --chains 2
run --atoms 200000 --chains 4
run --atoms 2000000 --steps 6000
run --atoms 2000000 --steps 6000 --chains 4
run --atoms 4000000 --steps 3000 --chains 4
