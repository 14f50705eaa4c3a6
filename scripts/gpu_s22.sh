#!/bin/bash
O=gpurun_out/s22; mkdir -p $O
CONSTV_DRUDE_SPT=2 timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude_spt2.log 2>&1; echo "pytest spt2 rc=$?" >> $O/status.txt
CONSTV_DRUDE_SPT=3 timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude_spt3.log 2>&1; echo "pytest spt3 rc=$?" >> $O/status.txt
CONSTV_DRUDE_SPT=2 CONSTV_DRUDE_MINB=6 timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude_spt2_minb6.log 2>&1; echo "pytest spt2 minb6 rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude > $O/ab_drude.jsonl 2> $O/ab_drude.err; echo "drude rc=$?" >> $O/status.txt
timeout 300 python scripts/ab_variants.py settle > $O/ab_settle_check.jsonl 2> $O/ab_settle.err; echo "settle rc=$?" >> $O/status.txt
echo done >> $O/status.txt
