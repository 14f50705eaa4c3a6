#!/bin/bash
# round 2 session 42: the final build: GPU suite, smoke, the driver's command (with the rigid-water legs of plugin_path)
O=gpurun_out/s42; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
/usr/bin/time -v -o $O/time_k20.txt timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_k20.json 2> $O/bench_ref_k20.err; echo "benchref rc=$?" >> $O/status.txt
echo done >> $O/status.txt
