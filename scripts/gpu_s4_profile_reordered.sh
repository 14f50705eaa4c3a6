#!/bin/bash
# ncu summary of the reordered-slot Langevin kernel (the state behind OpenMM after reorderAtoms)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --reordered"
$CMD > gpurun_out/reord_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_indexed -s 14 -c 2 -f -o gpurun_out/s4_prof_indexed $CMD > gpurun_out/reord_ncu_full.log 2>&1
tail -1 gpurun_out/reord_ncu_full.log
$CMD > /dev/null 2>&1 &&
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:constv_fused_step_indexed -s 12 -c 6 --csv --log-file gpurun_out/s4_dram_indexed.csv $CMD > /dev/null 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/s4_dram_indexed.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
