#!/bin/bash
# round 2 session 71: the ensemble of rigid-water slabs with bulk-staged tiles (constv_settle_step_bulk_batched_kernel): parity, then the flagged run
O=gpurun_out/s71; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_settle.py -m gpu -x -q -k "batched or guard" > $O/pytest_batched.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
V="--steps 600 --warmup 20 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
timeout 200 python bench.py --rigid-water fused --ensemble 64 --atoms 100000 --variant 7 $V > $O/bench_ensemble_rigid_cp_async.json 2>> $O/err.txt; echo "cp.async rc=$?" >> $O/status.txt
timeout 200 python bench.py --rigid-water fused --ensemble 64 --atoms 100000 $V > $O/bench_ensemble_rigid_bulk.json 2>> $O/err.txt; echo "bulk rc=$?" >> $O/status.txt
echo done >> $O/status.txt
