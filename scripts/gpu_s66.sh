#!/bin/bash
# round 2 session 66: ncu launch lists of the final build: the headline command and the rigid-water leg (each after a plain run that exited 0)
set -u
O=gpurun_out/s66; mkdir -p $O
CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?" >> $O/status.txt
CMD2="python bench.py --rigid-water fused --chains 6 --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3"
$CMD2 > $O/plain_rigid.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_rigid_water.csv $CMD2 > $O/ncu_launches_rigid.log 2>&1
echo "launch list rigid rc=$?" >> $O/status.txt
echo done >> $O/status.txt
