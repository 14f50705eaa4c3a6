#!/bin/bash
# round 2, GPU session 7: the two-ring TMA pipeline of the rigid-water step
O=gpurun_out/s7; mkdir -p $O
for g in 4 3 6; do
CONSTV_WS_GROUPS=$g timeout 300 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q > $O/pytest_ws_g${g}.log 2>&1; echo "pytest_ws g$g rc=$?" >> $O/status.txt
done
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-ref-gpu > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
