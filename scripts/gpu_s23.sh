#!/bin/bash
# round 2, GPU session 23: ncu --set full of the balanced Drude kernel (SPT=2); the ncu pass follows a plain run of the identical command
set -u
O=gpurun_out/s23; mkdir -p $O
export CONSTV_DRUDE_SPT=2
CMD="python bench.py --drude --chains 1 --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_drude_step_balanced_kernel -s 20 -c 3 -f -o $O/prof_drude_balanced $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?" >> $O/status.txt
echo done >> $O/status.txt
