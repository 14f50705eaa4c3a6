#!/bin/bash
# round 2 session 60: bulk-staged rigid-water step with the Philox draws made while the bulk copies fly (CONSTV_SETTLE_EARLY=1)
O=gpurun_out/s60; mkdir -p $O
CONSTV_SETTLE_EARLY=1 timeout 600 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q -k bulk_stage_in > $O/pytest_early.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk_early > $O/ab_early.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk_early 1000000 mixed > $O/ab_early_mixed.jsonl 2>> $O/ab.err; echo "ab mixed rc=$?" >> $O/status.txt
echo done >> $O/status.txt
