#!/bin/bash
# round 2, GPU session 12: the LDG-staged rigid-water kernel with cp.async stage-in and early Philox draws
O=gpurun_out/s12; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_settle.py tests/test_gpu_settle_ws.py tests/test_example_system.py tests/test_plugin_cpp.py -m gpu -x -q > $O/pytest_settle.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
echo done >> $O/status.txt
