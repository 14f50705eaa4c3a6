#!/bin/bash
# round 2 session 34: the driver-style runs of the final state of the code on ONE box, plus the flagged configurations
O=gpurun_out/s34; mkdir -p $O
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_k20.json 2> $O/bench_ref_k20.err; echo "benchref rc=$?" >> $O/status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 900 python bench.py --skip-cpu > $O/bench_default.json 2> $O/bench_default.err; echo "benchdef rc=$?" >> $O/status.txt
V="--steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu"
timeout 300 python bench.py --rigid-water fused --chains 6 $V --skip-ref-gpu > $O/bench_rigid_water.json 2>> $O/var.err; echo "rigid rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --chains 1 $V --skip-ref-gpu > $O/bench_rigid_water_single_launch.json 2>> $O/var.err; echo "rigid1 rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --chains 6 $V > $O/bench_drude.json 2>> $O/var.err; echo "drude rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --chains 1 $V --skip-ref-gpu > $O/bench_drude_single_launch.json 2>> $O/var.err; echo "drude1 rc=$?" >> $O/status.txt
timeout 300 python bench.py --reordered $V --skip-ref-gpu > $O/bench_reordered.json 2>> $O/var.err; echo "reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --reordered $V --skip-ref-gpu > $O/bench_rigid_water_reordered.json 2>> $O/var.err; echo "rigid reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --drude --reordered $V --skip-ref-gpu > $O/bench_drude_reordered.json 2>> $O/var.err; echo "drude reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --ensemble 64 --atoms 100000 --rigid-water fused $V --skip-ref-gpu > $O/bench_ensemble_rigid_water.json 2>> $O/var.err; echo "ensemble rigid rc=$?" >> $O/status.txt
timeout 300 python bench.py --atoms 4000000 --rigid-water fused --chains 2 --steps 1000 --warmup 30 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/bench_rigid_water_4M.json 2>> $O/var.err; echo "rigid 4M rc=$?" >> $O/status.txt
timeout 300 python bench.py --atoms 4000000 --drude --chains 2 --steps 1000 --warmup 30 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/bench_drude_4M.json 2>> $O/var.err; echo "drude 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
