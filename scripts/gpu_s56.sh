#!/bin/bash
# round 2 session 56: the full GPU suite with the bulk stage-in as the automatic choice; the rigid-water bench legs
O=gpurun_out/s56; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
V="--steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
timeout 300 python bench.py --rigid-water fused --chains 6 $V > $O/bench_rigid_water.json 2>> $O/var.err; echo "rigid rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --chains 1 $V > $O/bench_rigid_water_single_launch.json 2>> $O/var.err; echo "rigid1 rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --reordered $V > $O/bench_rigid_water_reordered.json 2>> $O/var.err; echo "rigid reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --atoms 4000000 --chains 3 $V > $O/bench_rigid_water_4M.json 2>> $O/var.err; echo "rigid 4M rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision mixed --rigid-water fused --chains 6 $V > $O/bench_mixed_rigid_water.json 2>> $O/var.err; echo "mixed rigid rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision mixed --rigid-water fused --chains 1 $V > $O/bench_mixed_rigid_water_single.json 2>> $O/var.err; echo "mixed rigid1 rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision mixed --rigid-water fused --reordered $V > $O/bench_mixed_rigid_water_reordered.json 2>> $O/var.err; echo "mixed rigid reordered rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision double --rigid-water fused --chains 6 $V > $O/bench_double_rigid_water.json 2>> $O/var.err; echo "double rigid rc=$?" >> $O/status.txt
for m in "--rigid-water" "--rigid-water --reordered" "--rigid-water --precision mixed" "--rigid-water --precision mixed --reordered"; do timeout 300 openmm_constv_b200/plugin/bin/BenchPluginPath --atoms 1000000 --steps 300 $m; done > $O/plugin_path_rigid.jsonl 2> $O/plugin.err; echo "plugin rc=$?" >> $O/status.txt
echo done >> $O/status.txt
