#!/bin/bash
O=gpurun_out/s35; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
V="--steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
timeout 300 python bench.py --rigid-water fused --reordered $V > $O/bench_rigid_water_reordered.json 2>> $O/var.err; echo "rigid reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --reordered $V > $O/bench_drude_reordered.json 2>> $O/var.err; echo "drude reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --chains 1 $V > $O/bench_rigid_water_single.json 2>> $O/var.err; echo "rigid single rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --chains 1 $V > $O/bench_drude_single.json 2>> $O/var.err; echo "drude single rc=$?" >> $O/status.txt
echo done >> $O/status.txt
