#!/bin/bash
O=gpurun_out/s33; mkdir -p $O
CONSTV_STAGED_MINB=9 timeout 600 python -m pytest tests/test_gpu_drude.py tests/test_gpu_settle.py -m gpu -x -q > $O/pytest_minb9.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle_minb > $O/ab_settle_minb.jsonl 2> $O/ab_settle_minb.err; echo "settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude_minb > $O/ab_drude_minb.jsonl 2> $O/ab_drude_minb.err; echo "drude rc=$?" >> $O/status.txt
echo done >> $O/status.txt
