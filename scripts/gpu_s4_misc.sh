#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20000 --warmup 500 2>gpurun_out/s4_scale2.err | tail -1 > gpurun_out/s4_scale_1M_2.json
python -c "import json; d=json.load(open('gpurun_out/s4_scale_1M_2.json')); print('2 GPUs', d['value'], d['ms_per_step'], d['config']['chains'])" || tail -5 gpurun_out/s4_scale2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 --warmup 3 2>>gpurun_out/s4_scale2.err | tail -1
