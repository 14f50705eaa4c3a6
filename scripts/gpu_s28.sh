#!/bin/bash
O=gpurun_out/s28; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude > $O/ab_drude.jsonl 2> $O/ab_drude.err; echo "drude rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude 4000000 > $O/ab_drude_4M.jsonl 2> $O/ab_drude_4M.err; echo "drude 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
