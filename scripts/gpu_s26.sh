#!/bin/bash
O=gpurun_out/s26; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench k20 rc=$?" >> $O/status.txt
EXE=openmm_constv_b200/plugin/bin/BenchPluginPath
for pf in 0 1; do
  CONSTV_PREFETCH_OWN=$pf timeout 300 $EXE --atoms 1000000 --steps 300 --precision single > $O/plugin_identity_prefetch$pf.json 2>&1; echo "plugin pf$pf rc=$?" >> $O/status.txt
done
timeout 900 python scripts/ab_variants.py drude 4000000 > $O/ab_drude_4M.jsonl 2> $O/ab_drude_4M.err; echo "drude 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
