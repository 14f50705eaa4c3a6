#!/bin/bash
O=gpurun_out/s15; mkdir -p $O
timeout 600 python scripts/ab_variants.py reordered > $O/ab_reordered.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py reordered 4000000 >> $O/ab_reordered.jsonl 2>> $O/ab.err; echo "ab4m rc=$?" >> $O/status.txt
CONSTV_FUSED_ZERO=5 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
CONSTV_FUSED_ZERO=5 timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered_prefetch.json 2>> $O/ab.err
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered.json 2>> $O/ab.err
echo done >> $O/status.txt
