#!/bin/bash
# round 2, final code: the N = $1 arm exactly as the driver launches it
N=$1; O=gpurun_out/n${N}b; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?" >> $O/status.txt
echo done >> $O/status.txt
