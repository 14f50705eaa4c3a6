#!/bin/bash
# round 2 session 65 (8 GPUs): the N = 8 arm exactly as the driver launches it, the last build and protocol
O=gpurun_out/s65; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
