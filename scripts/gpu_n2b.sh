#!/bin/bash
# round 2 session 30: the N = 2 arm exactly as the driver launches it (new rigid-water / Drude legs included)
O=gpurun_out/n2b; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
