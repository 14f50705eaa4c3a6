#!/bin/bash
# Run on the GPU box via gpurun: launch-shape sweep of the fused kernel, then the two ncu passes of
# B200_PROFILING.md on a short eager bench command (after the same command exited 0 without ncu).
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for ilp in 1 2 4; do for ctas in 2 3 4 6 8; do
  python bench.py --steps 6000 --warmup 300 --skip-cpu --skip-e2e --pairs-per-thread $ilp --ctas-per-sm $ctas 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'ilp':$ilp,'ctas':$ctas,'us':d['ms_per_step']*1e3,'GBs':d['roofline']['achieved'],'frac':d['roofline']['frac']}))" \
    | tee -a gpurun_out/sweep.jsonl
done; done
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 14 -c 3 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
