#!/bin/bash
O=gpurun_out/s25; mkdir -p $O
for shape in "64 2" "64 4" "32 4" "128 2"; do
  set -- $shape
  CONSTV_DRUDE_BLOCK=$1 CONSTV_DRUDE_SPT=$2 timeout 600 python -m pytest tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_drude_$1x$2.log 2>&1; echo "pytest $1x$2 rc=$?" >> $O/status.txt
done
timeout 1200 python scripts/ab_variants.py drude > $O/ab_drude.jsonl 2> $O/ab_drude.err; echo "drude rc=$?" >> $O/status.txt
echo done >> $O/status.txt
