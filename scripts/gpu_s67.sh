#!/bin/bash
# round 2 session 67: the driver's command once more after adding the `summary` key (last key of the line)
O=gpurun_out/s67; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --chains 6 --steps 20 --warmup 5 --skip-also --skip-cpu --skip-e2e --skip-ref-gpu > $O/bench_rigid_k20.json 2>> $O/bench_k20.err; echo "rigid rc=$?" >> $O/status.txt
echo done >> $O/status.txt
