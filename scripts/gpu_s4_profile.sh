#!/bin/bash
# session 4 ncu evidence for the bench default (2 chains).  Each ncu pass follows a plain run of the
# identical command that exited 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu"
$CMD > gpurun_out/s4_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv $CMD > gpurun_out/s4_ncu_launches.log 2>&1
$CMD > gpurun_out/s4_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 28 -c 4 -f -o gpurun_out/s4_prof_fused $CMD > gpurun_out/s4_ncu_full.log 2>&1
$CMD > gpurun_out/s4_plain3.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
    -k regex:constv_fused_step_kernel -s 24 -c 24 --csv --log-file gpurun_out/s4_dram_steady.csv $CMD > gpurun_out/s4_ncu_steady.log 2>&1
tail -2 gpurun_out/s4_ncu_full.log; tail -4 gpurun_out/s4_dram_steady.csv
python bench.py 2>gpurun_out/s4_bench_err.log | tail -1 > gpurun_out/s4_bench_default.json; cat gpurun_out/s4_bench_default.json
python bench.py --chains 1 --skip-cpu --skip-e2e --skip-ref-gpu 2>>gpurun_out/s4_bench_err.log | tail -1 > gpurun_out/s4_bench_chains1.json; cat gpurun_out/s4_bench_chains1.json
python bench.py --impl reference --steps 100 --warmup 5 | tail -1 > gpurun_out/s4_bench_reference.json; cat gpurun_out/s4_bench_reference.json
for c in 1 2 4; do
python bench.py --ensemble 8 --atoms 100000 --chains $c --steps 20000 --warmup 500 2>>gpurun_out/s4_bench_err.log | tail -1 > gpurun_out/s4_bench_ensemble_c$c.json
python -c "import json; d=json.load(open('gpurun_out/s4_bench_ensemble_c$c.json')); print('ensemble chains $c', d['ms_per_step']*1e3, d['roofline']['frac'])"
done
