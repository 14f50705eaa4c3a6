#!/bin/bash
# round 2, GPU session 2: where to put the Philox draw and the zeroing stores of the fused step; the plugin path
O=gpurun_out/s2; mkdir -p $O
timeout 900 python -m pytest tests/test_example_system.py tests/test_gpu_philox.py -m gpu -x -q > $O/pytest_new.log 2>&1; echo "pytest_new rc=$?" >> $O/status.txt
for early in 0 1; do for zero in 0 1 2; do
  CONSTV_FUSED_EARLY=$early CONSTV_FUSED_ZERO=$zero timeout 300 python bench.py --steps 2000 --warmup 50 --skip-cpu --skip-e2e --skip-ref-gpu --only single > $O/var_e${early}_z${zero}.json 2>> $O/var.err
  CONSTV_FUSED_EARLY=$early CONSTV_FUSED_ZERO=$zero timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_e${early}_z${zero}.json 2>> $O/var.err
done; done
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered > $O/plugin_reordered.json 2>> $O/var.err
CONSTV_CACHE_IMAGE_CHARGE=0 timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 > $O/plugin_nocache.json 2>> $O/var.err
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --atoms 2000000 > $O/plugin_2m.json 2>> $O/var.err
for p in single mixed double; do timeout 600 openmm_constv_b200/plugin/bin/TestCudaConstVLangevinIntegrator $p >> $O/cpptest.log 2>&1; done
echo done >> $O/status.txt
