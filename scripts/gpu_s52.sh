#!/bin/bash
# round 2 session 52: constv_step_host_submit_state (pipelined read-back of posq / velm): parity tests and the e2e legs
O=gpurun_out/s52; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or pipelined" > $O/pytest_host.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 20 --warmup 5 --skip-also --skip-cpu --skip-ref-gpu > $O/bench_e2e.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
echo done >> $O/status.txt
