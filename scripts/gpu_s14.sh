#!/bin/bash
O=gpurun_out/s14; mkdir -p $O
timeout 600 python scripts/ab_variants.py reordered > $O/ab_reordered.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_variants.py reordered 4000000 >> $O/ab_reordered.jsonl 2>> $O/ab.err; echo "ab4m rc=$?" >> $O/status.txt
CONSTV_INDEXED_MINB=8 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -x -q -k "reorder or perm or index" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
echo done >> $O/status.txt
