#!/bin/bash
# round 2 session 63: the final build as the driver runs it: GPU suite, smoke, reference arm, the driver's command, the default command
O=gpurun_out/s63; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_k20.json 2> $O/bench_ref_k20.err; echo "benchref rc=$?" >> $O/status.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err ) 2> $O/bench_k20.time; echo "bench20 rc=$?" >> $O/status.txt
timeout 900 python bench.py --skip-cpu > $O/bench_default.json 2> $O/bench_default.err; echo "benchdef rc=$?" >> $O/status.txt
echo done >> $O/status.txt
