#!/bin/bash
set -u
mkdir -p gpurun_out
run() { python bench.py --steps 1000 --warmup 100 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err.log | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4),'clocks':d['clocks']}))" \
    | tee -a gpurun_out/sweep5.jsonl || tail -3 gpurun_out/err.log; }
: > gpurun_out/sweep5.jsonl
run --atoms 16000000
run --atoms 16000000 --no-cache-image-charge
run --atoms 4000000
run --atoms 4000000 --no-cache-image-charge
A="python bench.py --steps 6 --warmup 3 --no-graph --skip-cpu --skip-e2e --atoms 4000000 --replicas 3"
$A > gpurun_out/plainA.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 5 -c 2 -f -o gpurun_out/prof_cached $A > gpurun_out/ncuA.log 2>&1
B="$A --no-cache-image-charge"
$B > gpurun_out/plainB.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 5 -c 2 -f -o gpurun_out/prof_nocache $B > gpurun_out/ncuB.log 2>&1
tail -2 gpurun_out/ncuA.log gpurun_out/ncuB.log
