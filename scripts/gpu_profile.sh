#!/bin/bash
# ncu evidence for the default fused kernel (run under gpurun).  Each ncu pass follows a plain run of
# the identical command that exited 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 14 -c 3 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
    -k regex:constv_fused_step_kernel -s 12 -c 12 --csv --log-file gpurun_out/dram_steady.csv $CMD > gpurun_out/ncu_steady.log 2>&1
tail -2 gpurun_out/ncu_full.log; tail -4 gpurun_out/dram_steady.csv
python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_default.json; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 100 --warmup 5 | tail -1 > gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.json
