#!/bin/bash
# round 2 session 49: host topology and pinned H2D bandwidth per memory node (scripts/numa_probe.py); GPU suite on the rebuilt library
O=gpurun_out/s49; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | head -40 > $O/lscpu.txt 2>&1
timeout 300 python scripts/numa_probe.py > $O/numa_probe.jsonl 2> $O/numa_probe.err; echo "probe rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
echo done >> $O/status.txt
