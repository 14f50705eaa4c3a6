#!/bin/bash
O=gpurun_out/s27; mkdir -p $O
CONSTV_SETTLE_EARLY=2 timeout 600 python -m pytest tests/test_gpu_settle.py tests/test_example_system.py -m gpu -x -q > $O/pytest_settle_cpasync.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4M.jsonl 2> $O/ab_settle_4M.err; echo "settle 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
