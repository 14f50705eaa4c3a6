#!/bin/bash
O=gpurun_out/s41; mkdir -p $O
EXE=openmm_constv_b200/plugin/bin/BenchPluginPath
for prec in single mixed; do
  for mode in "" "--reordered"; do
    tag=$prec$(echo $mode | tr -d ' -')
    timeout 300 $EXE --atoms 1000000 --steps 300 --precision $prec --rigid-water $mode > $O/plugin_rigid_$tag.json 2> $O/plugin_rigid_$tag.err; echo "$tag rc=$?" >> $O/status.txt
  done
done
timeout 300 $EXE --atoms 1000000 --steps 300 --precision mixed > $O/plugin_mixed.json 2>&1; echo "mixed rc=$?" >> $O/status.txt
timeout 300 $EXE --atoms 1000000 --steps 300 --precision mixed --reordered > $O/plugin_mixedreordered.json 2>&1; echo "mixed reordered rc=$?" >> $O/status.txt
echo done >> $O/status.txt
