#!/bin/bash
# round 2 session 31: ncu evidence for the default staged kernels (rigid water: cp.async stage-in; Drude: balanced kernel).
# Each ncu pass follows a plain run of the identical command that exited 0 (B200_PROFILING.md).
set -u
O=gpurun_out/s31; mkdir -p $O
COMMON="--steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3 --chains 1"
for kind in rigid drude; do
  if [ $kind = rigid ]; then FLAG="--rigid-water fused"; K=constv_settle_step_tiled_kernel; else FLAG="--drude"; K=constv_drude_step_balanced_kernel; fi
  CMD="python bench.py $FLAG $COMMON"
  $CMD > $O/plain_$kind.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 3 -f -o $O/prof_$kind $CMD > $O/ncu_full_$kind.log 2>&1
  echo "$kind full rc=$?" >> $O/status.txt
  $CMD > $O/plain2_$kind.log 2>&1 &&
  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      -k regex:$K -s 20 -c 24 --csv --log-file $O/dram_steady_$kind.csv $CMD > $O/ncu_steady_$kind.log 2>&1
  echo "$kind steady rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
