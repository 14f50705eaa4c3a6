#!/bin/bash
O=gpurun_out/s17; mkdir -p $O
for rep in 1 2; do for c in 2 3 4 8; do
timeout 300 python bench.py --steps 20 --warmup 5 --chains $c --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/k20_c${c}_r${rep}.json 2>> $O/err.log
done; done
echo done >> $O/status.txt
