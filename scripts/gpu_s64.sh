#!/bin/bash
# round 2 session 64 (2 GPUs): the N = 2 arm exactly as the driver launches it, with the last build
O=gpurun_out/s64; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref n2 rc=$?" >> $O/status.txt
echo done >> $O/status.txt
