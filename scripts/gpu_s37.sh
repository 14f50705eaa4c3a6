#!/bin/bash
O=gpurun_out/s37; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
for m in settle_reordered drude_reordered drude_chains; do
  timeout 600 python scripts/ab_variants.py $m > $O/ab_$m.jsonl 2> $O/ab_$m.err; echo "$m rc=$?" >> $O/status.txt
done
timeout 600 python scripts/ab_variants.py drude 4000000 > $O/ab_drude_4M.jsonl 2> $O/ab_drude_4M.err; echo "drude 4M rc=$?" >> $O/status.txt
echo done >> $O/status.txt
