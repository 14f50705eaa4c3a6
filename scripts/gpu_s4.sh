#!/bin/bash
# round 2, GPU session 4: parity of the TMA-pipelined rigid-water kernel, its A/B against the LDG-staged kernel, plugin path
O=gpurun_out/s4; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q > $O/pytest_ws.log 2>&1; wsrc=$?; echo "pytest_ws rc=$wsrc" >> $O/status.txt
if [ $wsrc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_settle.py tests/test_example_system.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_settle.log 2>&1; echo "pytest_settle rc=$?" >> $O/status.txt
  timeout 600 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
  timeout 600 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
fi
for early in 0 1; do
  CONSTV_FUSED_EARLY=$early CONSTV_FUSED_ZERO=$early timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_e${early}.json 2>> $O/var.err
done
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered.json 2>> $O/var.err
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
