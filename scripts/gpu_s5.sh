#!/bin/bash
# round 2, GPU session 5: ncu of the TMA-pipelined rigid-water kernel and of the LDG-staged one
O=gpurun_out/s5; mkdir -p $O
CMD="python bench.py --rigid-water fused --chains 1 --steps 20 --warmup 5 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu --repeats 3"
$CMD > $O/plain_ws.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:settle_step_ws -s 30 -c 3 -o $O/ws $CMD > $O/ncu_ws.log 2>&1
echo "ncu ws rc=$?" >> $O/status.txt
CONSTV_SETTLE_WS=0 $CMD > $O/plain_ldg.log 2>&1 && CONSTV_SETTLE_WS=0 ncu --set full --clock-control none --import-source on -k regex:settle_step_tiled -s 30 -c 3 -o $O/ldg $CMD > $O/ncu_ldg.log 2>&1
echo "ncu ldg rc=$?" >> $O/status.txt
echo done >> $O/status.txt
