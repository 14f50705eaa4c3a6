#!/bin/bash
O=gpurun_out/s47; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_plugin_cpp.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py reordered > $O/ab_reordered.jsonl 2> $O/ab_reordered.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py reordered 4000000 > $O/ab_reordered_4M.jsonl 2> $O/ab_reordered_4M.err; echo "ab 4M rc=$?" >> $O/status.txt
EXE=openmm_constv_b200/plugin/bin/BenchPluginPath
for t in 0 1; do
  CONSTV_IMAGE_TABLES=$t timeout 300 $EXE --atoms 1000000 --steps 300 --precision single --reordered > $O/plugin_reordered_tables$t.json 2>&1; echo "plugin $t rc=$?" >> $O/status.txt
  CONSTV_IMAGE_TABLES=$t timeout 300 $EXE --atoms 1000000 --steps 300 --precision single --reordered --rigid-water > $O/plugin_rigid_reordered_tables$t.json 2>&1; echo "plugin rigid $t rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
