#!/bin/bash
# session 4: weak scaling with two chains at 8 GPUs, and the 16M slab (2M atoms per GPU)
set -u
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50000 --warmup 1000 --skip-e2e 2>gpurun_out/s4_scale$N.err | tail -1 > gpurun_out/s4_scale_1M_$N.json
python -c "import json; d=json.load(open('gpurun_out/s4_scale_1M_$N.json')); print('$N GPUs 1M/GPU', d['value'], d['ms_per_step'], d['roofline']['frac'])" || tail -5 gpurun_out/s4_scale$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --atoms $((16000000/N)) --steps 20000 --warmup 500 --skip-e2e 2>>gpurun_out/s4_scale$N.err | tail -1 > gpurun_out/s4_scale_16M_$N.json
python -c "import json; d=json.load(open('gpurun_out/s4_scale_16M_$N.json')); print('$N GPUs 16M slab', d['value'], d['ms_per_step'], d['roofline']['frac'])" || tail -5 gpurun_out/s4_scale$N.err
