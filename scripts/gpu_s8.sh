#!/bin/bash
O=gpurun_out/s8; mkdir -p $O
timeout 300 python scripts/ws_repro.py 1000000 30 4 > $O/repro_g4.log 2>&1; echo "repro g4 rc=$?" >> $O/status.txt
timeout 300 python scripts/ws_repro.py 1000000 30 3,5,6 > $O/repro_g356.log 2>&1; echo "repro g356 rc=$?" >> $O/status.txt
timeout 600 compute-sanitizer --tool memcheck python scripts/ws_repro.py 400000 3 4 > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/status.txt
echo done >> $O/status.txt
