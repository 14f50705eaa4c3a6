#!/bin/bash
# round 2, GPU session 1: tests, the bench at the driver's arguments, kernel variants of the single launch
O=gpurun_out/s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 600 python bench.py --skip-cpu --skip-ref-gpu > $O/bench_default.json 2> $O/bench_default.err; echo "benchdef rc=$?" >> $O/status.txt
for early in 1 0; do for blk in 256 128; do
  CONSTV_FUSED_EARLY=$early CONSTV_FUSED_BLOCK=$blk timeout 300 python bench.py --steps 2000 --warmup 50 --skip-cpu --skip-e2e --skip-ref-gpu --only single > $O/var_e${early}_b${blk}.json 2>> $O/var.err
  CONSTV_FUSED_EARLY=$early CONSTV_FUSED_BLOCK=$blk timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_e${early}_b${blk}.json 2>> $O/var.err
done; done
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered.json 2>> $O/var.err
CONSTV_CACHE_IMAGE_CHARGE=0 timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_nocache.json 2>> $O/var.err
timeout 300 python bench.py --reordered --steps 2000 --warmup 50 --skip-cpu --skip-e2e --skip-ref-gpu --skip-also > $O/reordered.json 2>> $O/var.err
echo done >> $O/status.txt
