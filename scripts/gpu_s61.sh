#!/bin/bash
# round 2 session 61: full GPU suite with the final defaults; mixed-precision rigid-water legs (early draws on by default there)
O=gpurun_out/s61; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
V="--steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
timeout 300 python bench.py --precision mixed --rigid-water fused --chains 3 $V > $O/bench_mixed_rigid_water_3chains.json 2>> $O/var.err; echo "mixed rigid3 rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision mixed --rigid-water fused --chains 6 $V > $O/bench_mixed_rigid_water.json 2>> $O/var.err; echo "mixed rigid6 rc=$?" >> $O/status.txt
timeout 300 python bench.py --precision mixed --rigid-water fused --chains 1 $V > $O/bench_mixed_rigid_water_single.json 2>> $O/var.err; echo "mixed rigid1 rc=$?" >> $O/status.txt
for m in "--rigid-water --precision mixed"; do timeout 300 openmm_constv_b200/plugin/bin/BenchPluginPath --atoms 1000000 --steps 300 $m; done > $O/plugin_path_rigid_mixed.jsonl 2> $O/plugin.err; echo "plugin rc=$?" >> $O/status.txt
echo done >> $O/status.txt
