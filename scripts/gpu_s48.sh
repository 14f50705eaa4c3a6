#!/bin/bash
# round 2 session 48: the final build once more: GPU suite, smoke, the driver's command, the default command, the reordered flagged runs
O=gpurun_out/s48; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_k20.json 2> $O/bench_ref_k20.err; echo "benchref rc=$?" >> $O/status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 900 python bench.py --skip-cpu > $O/bench_default.json 2> $O/bench_default.err; echo "benchdef rc=$?" >> $O/status.txt
V="--steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
timeout 300 python bench.py --reordered $V > $O/bench_reordered.json 2>> $O/var.err; echo "reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --reordered $V > $O/bench_rigid_water_reordered.json 2>> $O/var.err; echo "rigid reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --reordered $V > $O/bench_drude_reordered.json 2>> $O/var.err; echo "drude reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --chains 6 $V > $O/bench_drude.json 2>> $O/var.err; echo "drude rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --chains 1 $V > $O/bench_drude_single_launch.json 2>> $O/var.err; echo "drude1 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
