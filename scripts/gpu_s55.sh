#!/bin/bash
# round 2 session 55: bulk stage-in on reordered slots, the automatic rule, constv_last_kernel: parity, then A/B
O=gpurun_out/s55; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_settle_ws.py tests/test_gpu_settle.py tests/test_example_system.py tests/test_plugin_cpp.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk > $O/ab_settle_bulk.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk_reordered > $O/ab_settle_bulk_reordered.jsonl 2>> $O/ab.err; echo "ab reordered rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py settle_bulk_reordered 1000000 mixed > $O/ab_settle_bulk_reordered_mixed.jsonl 2>> $O/ab.err; echo "ab reordered mixed rc=$?" >> $O/status.txt
echo done >> $O/status.txt
