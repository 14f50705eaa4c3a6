#!/bin/bash
# round 2, GPU session 3: parity of the TMA-pipelined rigid-water kernel, in-process A/B of variants, plugin path
O=gpurun_out/s3; mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q > $O/pytest_ws.log 2>&1; wsrc=$?; echo "pytest_ws rc=$wsrc" >> $O/status.txt
if [ $wsrc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_settle.py tests/test_example_system.py -m gpu -x -q > $O/pytest_settle.log 2>&1; echo "pytest_settle rc=$?" >> $O/status.txt
  timeout 600 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
  timeout 600 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
fi
timeout 900 python scripts/ab_variants.py fused > $O/ab_fused.jsonl 2> $O/ab_fused.err; echo "ab_fused rc=$?" >> $O/status.txt
for early in 0 1; do
  CONSTV_FUSED_EARLY=$early timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_e${early}.json 2>> $O/var.err
  CONSTV_FUSED_EARLY=$early timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 >> $O/plugin_e${early}.json 2>> $O/var.err
done
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered.json 2>> $O/var.err
echo done >> $O/status.txt
