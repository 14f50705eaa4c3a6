#!/bin/bash
# round 2, GPU session 6: shapes of the TMA-pipelined rigid-water kernel (consumer groups x straight-line compute)
O=gpurun_out/s6; mkdir -p $O
for g in 7 4; do for ilp in 0 1; do
CONSTV_WS_GROUPS=$g CONSTV_WS_ILP=$ilp timeout 300 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q > $O/pytest_ws_g${g}_i${ilp}.log 2>&1; echo "pytest_ws g$g ilp$ilp rc=$?" >> $O/status.txt
done; done
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
echo done >> $O/status.txt
