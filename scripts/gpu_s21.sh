#!/bin/bash
O=gpurun_out/s21; mkdir -p $O
CONSTV_PREFETCH_OWN=1 CONSTV_PREFETCH_AHEAD=3 timeout 600 python -m pytest tests/test_gpu_settle.py tests/test_gpu_drude.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_prefetch.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py prefetch_settle > $O/ab_prefetch_settle.jsonl 2> $O/ab_prefetch_settle.err; echo "settle rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py prefetch_drude > $O/ab_prefetch_drude.jsonl 2> $O/ab_prefetch_drude.err; echo "drude rc=$?" >> $O/status.txt
timeout 400 python scripts/ab_variants.py prefetch_fused > $O/ab_prefetch_fused.jsonl 2> $O/ab_prefetch_fused.err; echo "fused rc=$?" >> $O/status.txt
echo done >> $O/status.txt
