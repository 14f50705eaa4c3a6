#!/bin/bash
# round 2: the N = 4 and N = 8 arms exactly as the driver launches them
O=gpurun_out/n8; mkdir -p $O
for n in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench n$n rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
