#!/bin/bash
# round 2 session 62: the region-boundary event nodes as side branches of the graph (BENCH_EVENT_BRANCH=1) against in-line (0)
O=gpurun_out/s62; mkdir -p $O
V="--steps 20 --warmup 5 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
for i in 1 2 3; do
BENCH_EVENT_BRANCH=0 timeout 300 python bench.py $V > $O/inline_$i.json 2>> $O/err.txt; echo "inline $i rc=$?" >> $O/status.txt
BENCH_EVENT_BRANCH=1 timeout 300 python bench.py $V > $O/branch_$i.json 2>> $O/err.txt; echo "branch $i rc=$?" >> $O/status.txt
done
BENCH_EVENT_BRANCH=1 timeout 300 python bench.py --steps 200 --warmup 5 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu --repeats 50 > $O/branch_k200.json 2>> $O/err.txt
BENCH_EVENT_BRANCH=0 timeout 300 python bench.py --steps 200 --warmup 5 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu --repeats 50 > $O/inline_k200.json 2>> $O/err.txt
echo done >> $O/status.txt
