#!/bin/bash
set -u
mkdir -p gpurun_out
true
true
: > gpurun_out/sweep_settle6.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err_settle.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); r=d.get('reference_gpu') or {}; print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4),'ref_us':round(r.get('ms_per_step',0)*1e3,3)}))" \
    | tee -a gpurun_out/sweep_settle6.jsonl || tail -5 gpurun_out/err_settle.log; }
run --rigid-water fused
run --rigid-water fused --variant 1
run --rigid-water split
run --rigid-water fused --atoms 4000000 --steps 3000
run --rigid-water fused --precision mixed
run --drude
run --drude --atoms 4000000 --steps 3000
run --drude --precision mixed
python bench.py --drude --steps 20000 --warmup 600 2>>gpurun_out/err_settle.log | tail -1 > gpurun_out/s4_bench_drude.json
python bench.py --rigid-water fused --steps 20000 --warmup 600 2>>gpurun_out/err_settle.log | tail -1 > gpurun_out/s4_bench_rigid_water_fused.json
python bench.py --rigid-water split --steps 20000 --warmup 600 2>>gpurun_out/err_settle.log | tail -1 > gpurun_out/s4_bench_rigid_water_split.json
