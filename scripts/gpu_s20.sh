#!/bin/bash
O=gpurun_out/s20; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
echo done >> $O/status.txt
