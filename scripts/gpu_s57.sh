#!/bin/bash
# round 2 session 57: ncu full set of the rigid-water step, bulk stage-in (variant 6) beside the cp.async-staged kernel (variant 7)
O=gpurun_out/s57; mkdir -p $O
V="--rigid-water fused --chains 1 --steps 12 --warmup 3 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu --no-graph --replicas 6"
timeout 300 python bench.py $V --variant 6 > $O/plain_bulk.json 2> $O/plain.err; echo "plain bulk rc=$?" >> $O/status.txt
timeout 300 python bench.py $V --variant 7 > $O/plain_cp_async.json 2>> $O/plain.err; echo "plain cp.async rc=$?" >> $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:constv_settle_step_bulk_kernel -s 8 -c 4 -o $O/settle_bulk -f python bench.py $V --variant 6 > $O/ncu_bulk.log 2>&1; echo "ncu bulk rc=$?" >> $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:constv_settle_step_tiled_kernel -s 8 -c 4 -o $O/settle_cp_async -f python bench.py $V --variant 7 > $O/ncu_cp_async.log 2>&1; echo "ncu cp.async rc=$?" >> $O/status.txt
ls -la $O >> $O/status.txt
echo done >> $O/status.txt
