#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_drude.py tests/test_plugin_cpp.py -m gpu -x -q > gpurun_out/pytest_gpu_drude.log 2>&1
tail -5 gpurun_out/pytest_gpu_drude.log
: > gpurun_out/sweep_drude.jsonl
run() { timeout 300 python bench.py --steps 12000 --warmup 600 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err_drude.log | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); r=d.get('reference_gpu') or {}; print(json.dumps({'args':'$*','us':round(d['ms_per_step']*1e3,3),'GBs':round(d['roofline']['achieved'],1),'frac':round(d['roofline']['frac'],4),'ref_us':round(r.get('ms_per_step',0)*1e3,3)}))" \
    | tee -a gpurun_out/sweep_drude.jsonl || tail -5 gpurun_out/err_drude.log; }
run --drude
run --drude --variant 1 --skip-ref-gpu
run --drude --atoms 4000000 --steps 3000
run --drude --precision mixed
python bench.py --drude --steps 20000 --warmup 600 2>>gpurun_out/err_drude.log | tail -1 > gpurun_out/s4_bench_drude.json
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --drude"
$CMD > gpurun_out/drude_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_drude_step -s 14 -c 2 -f -o gpurun_out/s4_prof_drude $CMD > gpurun_out/drude_ncu_full.log 2>&1
tail -2 gpurun_out/drude_ncu_full.log
