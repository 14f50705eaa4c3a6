#!/bin/bash
O=gpurun_out/s13; mkdir -p $O
for k in 1 2 3; do
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 >> $O/plugin_identity.json 2>> $O/var.err
timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 --reordered >> $O/plugin_reordered.json 2>> $O/var.err
done
CONSTV_FUSED_EARLY=0 CONSTV_FUSED_ZERO=0 timeout 120 openmm_constv_b200/plugin/bin/BenchPluginPath --steps 300 >> $O/plugin_identity_round1_order.json 2>> $O/var.err
timeout 300 python bench.py --reordered --steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/bench_reordered.json 2>> $O/var.err
echo done >> $O/status.txt
