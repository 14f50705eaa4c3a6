#!/bin/bash
# round 2 session 44: ncu of the rigid-water step in MIXED precision (the reference example's CudaPrecision); also verifies the aligned Drude stages
set -u
O=gpurun_out/s44; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude.log 2>&1; echo "pytest drude rc=$?" >> $O/status.txt
CMD="python bench.py --precision mixed --rigid-water fused --chains 1 --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu --skip-also --repeats 3"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constv_settle_step_tiled_kernel -s 20 -c 3 -f -o $O/prof_rigid_mixed $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?" >> $O/status.txt
echo done >> $O/status.txt
