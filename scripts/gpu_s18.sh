#!/bin/bash
# round 2, GPU session 10: the whole GPU suite, the bench at the driver's and at the default arguments, smoke
O=gpurun_out/s18; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_k20.json 2> $O/bench_ref_k20.err; echo "benchref rc=$?" >> $O/status.txt
timeout 600 python bench.py --skip-cpu --skip-ref-gpu > $O/bench_default.json 2> $O/bench_default.err; echo "benchdef rc=$?" >> $O/status.txt
timeout 300 python bench.py --rigid-water fused --chains 1 --steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/bench_rigid_water.json 2>> $O/var.err
timeout 300 python bench.py --drude --chains 1 --steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu > $O/bench_drude.json 2>> $O/var.err
timeout 300 python bench.py --reordered --steps 2000 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu > $O/bench_reordered.json 2>> $O/var.err
echo done >> $O/status.txt
