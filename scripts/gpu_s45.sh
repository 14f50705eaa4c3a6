#!/bin/bash
O=gpurun_out/s45; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_settle.py tests/test_gpu_drude.py tests/test_plugin_cpp.py tests/test_example_system.py -m gpu -x -q > $O/pytest_staged.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
for m in settle_reordered drude_reordered; do
  timeout 600 python scripts/ab_variants.py $m > $O/ab_$m.jsonl 2> $O/ab_$m.err; echo "$m rc=$?" >> $O/status.txt
  timeout 600 python scripts/ab_variants.py $m 1000000 mixed > $O/ab_${m}_mixed.jsonl 2> $O/ab_${m}_mixed.err; echo "$m mixed rc=$?" >> $O/status.txt
done
EXE=openmm_constv_b200/plugin/bin/BenchPluginPath
timeout 300 $EXE --atoms 1000000 --steps 300 --precision single --rigid-water --reordered > $O/plugin_rigid_singlereordered.json 2>&1; echo "plugin single rc=$?" >> $O/status.txt
timeout 300 $EXE --atoms 1000000 --steps 300 --precision mixed --rigid-water --reordered > $O/plugin_rigid_mixedreordered.json 2>&1; echo "plugin mixed rc=$?" >> $O/status.txt
echo done >> $O/status.txt
