#!/bin/bash
# round 2 session 51 (4 GPUs): pinned H2D bandwidth per rank alone and all together; then the e2e leg of the 4-rank bench
O=gpurun_out/s51; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc > $O/nproc.txt; cat /sys/devices/system/node/online >> $O/nproc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/h2d_concurrent_probe.py > $O/h2d_probe.jsonl 2> $O/h2d_probe.err; echo "probe rc=$?" >> $O/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 --skip-also --skip-cpu --skip-ref-gpu > $O/bench_n4_e2e.json 2> $O/bench_n4.err; echo "bench rc=$?" >> $O/status.txt
echo done >> $O/status.txt
