#!/bin/bash
O=gpurun_out/s36; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_settle.py tests/test_gpu_drude.py tests/test_gpu_settle_ws.py tests/test_example_system.py -m gpu -x -q > $O/pytest_staged.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
CONSTV_FUSED_ZERO=0 timeout 900 python -m pytest tests/test_gpu_settle.py tests/test_gpu_drude.py -m gpu -x -q > $O/pytest_staged_zero_late.log 2>&1; echo "pytest zero late rc=$?" >> $O/status.txt
for m in settle_zero drude_zero settle_zero_reordered drude_zero_reordered; do
  timeout 600 python scripts/ab_variants.py $m > $O/ab_$m.jsonl 2> $O/ab_$m.err; echo "$m rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
