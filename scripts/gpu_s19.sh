#!/bin/bash
# round 2, GPU session 11: ncu evidence for the bench default (2 chains).  Each ncu pass follows a plain run of the
# identical command that exited 0 (B200_PROFILING.md).
set -u
O=gpurun_out/s19; mkdir -p $O
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?" >> $O/status.txt
$CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_fused_step_kernel -s 40 -c 4 -f -o $O/prof_fused $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?" >> $O/status.txt
$CMD > $O/plain3.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
    -k regex:constv_fused_step_kernel -s 40 -c 24 --csv --log-file $O/dram_steady.csv $CMD > $O/ncu_steady.log 2>&1
echo "steady rc=$?" >> $O/status.txt
echo done >> $O/status.txt
