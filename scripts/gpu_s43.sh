#!/bin/bash
O=gpurun_out/s43; mkdir -p $O
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
t1=$(date +%s); echo "bench k20 wall seconds: $((t1-t0))" >> $O/status.txt
echo done >> $O/status.txt
