#!/bin/bash
# round 2 session 54: the rigid-water step with a bulk stage-in (kernel variant 6): parity, then A/B against the staged kernel
O=gpurun_out/s54; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_settle_ws.py -m gpu -x -q > $O/pytest_ws.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk > $O/ab_settle_bulk.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk 4000000 > $O/ab_settle_bulk_4M.jsonl 2>> $O/ab.err; echo "ab 4M rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk 1000000 mixed > $O/ab_settle_bulk_mixed.jsonl 2>> $O/ab.err; echo "ab mixed rc=$?" >> $O/status.txt
echo done >> $O/status.txt
