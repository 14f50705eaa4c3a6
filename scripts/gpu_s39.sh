#!/bin/bash
# round 2 session 39: mixed precision (what the reference's example runs: CudaPrecision mixed, nacl_constv.py:151) and double
O=gpurun_out/s39; mkdir -p $O
V="--steps 1500 --warmup 50 --skip-also --skip-e2e --skip-cpu --skip-ref-gpu"
for prec in mixed double; do
  timeout 300 python bench.py --precision $prec $V > $O/bench_${prec}.json 2>> $O/var.err; echo "$prec langevin rc=$?" >> $O/status.txt
  timeout 300 python bench.py --precision $prec --rigid-water fused --chains 6 $V > $O/bench_${prec}_rigid_water.json 2>> $O/var.err; echo "$prec rigid rc=$?" >> $O/status.txt
  timeout 300 python bench.py --precision $prec --rigid-water fused --chains 1 $V > $O/bench_${prec}_rigid_water_single.json 2>> $O/var.err; echo "$prec rigid1 rc=$?" >> $O/status.txt
  timeout 300 python bench.py --precision $prec --drude --chains 6 $V > $O/bench_${prec}_drude.json 2>> $O/var.err; echo "$prec drude rc=$?" >> $O/status.txt
  timeout 300 python bench.py --precision $prec --reordered $V > $O/bench_${prec}_reordered.json 2>> $O/var.err; echo "$prec reordered rc=$?" >> $O/status.txt
  timeout 300 python bench.py --precision $prec --rigid-water fused --reordered $V > $O/bench_${prec}_rigid_water_reordered.json 2>> $O/var.err; echo "$prec rigid reordered rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
