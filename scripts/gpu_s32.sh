#!/bin/bash
O=gpurun_out/s32; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
CONSTV_DRUDE_SPT=3 timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude_spt3.log 2>&1; echo "pytest spt3 rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude_chains > $O/ab_drude_chains.jsonl 2> $O/ab_drude_chains.err; echo "drude chains rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude 4000000 > $O/ab_drude_4M.jsonl 2> $O/ab_drude_4M.err; echo "drude 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
