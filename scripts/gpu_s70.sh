#!/bin/bash
# round 2 session 70 (4 GPUs): the N = 4 arm exactly as the driver launches it, the last build and protocol
O=gpurun_out/s70; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
