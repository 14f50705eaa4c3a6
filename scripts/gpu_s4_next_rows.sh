#!/bin/bash
# final evidence for the SURVEY 8f rows: bench JSON lines and ncu summaries of the staged rigid-water and Drude kernels
set -u
mkdir -p gpurun_out
python bench.py --drude --steps 20000 --warmup 600 2>gpurun_out/err_next.log | tail -1 > gpurun_out/s4_bench_drude.json
python bench.py --rigid-water fused --steps 20000 --warmup 600 2>>gpurun_out/err_next.log | tail -1 > gpurun_out/s4_bench_rigid_water_fused.json
python bench.py --rigid-water split --steps 20000 --warmup 600 2>>gpurun_out/err_next.log | tail -1 > gpurun_out/s4_bench_rigid_water_split.json
for f in drude rigid_water_fused rigid_water_split; do python -c "
import json; d=json.load(open('gpurun_out/s4_bench_$f.json')); print('$f', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],3), (d.get('reference_gpu') or {}).get('ms_per_step'))"; done
for mode in "--rigid-water fused" "--drude"; do
  tag=$(echo $mode | tr -d ' -')
  CMD="python bench.py --steps 18 --warmup 6 --no-graph --skip-cpu --skip-e2e --skip-ref-gpu $mode"
  $CMD > gpurun_out/next_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 14 -c 2 -f -o gpurun_out/s4_prof_$tag $CMD > gpurun_out/next_ncu_$tag.log 2>&1
  tail -1 gpurun_out/next_ncu_$tag.log
  $CMD > /dev/null 2>&1 &&
  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tiled_kernel -s 12 -c 6 --csv --log-file gpurun_out/s4_dram_$tag.csv $CMD > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time" gpurun_out/s4_dram_$tag.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
done
