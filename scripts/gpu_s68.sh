#!/bin/bash
# round 2 session 68: how much the swizzled stage of the Drude kernel is worth (a linear stage is what bulk copies would leave)
O=gpurun_out/s68; mkdir -p $O
cp openmm_constv_b200/libconstv_b200.so /tmp/lib_swz.so
for i in 1 2; do
cp /tmp/lib_swz.so openmm_constv_b200/libconstv_b200.so
AB_CHAINS=1,6 timeout 300 python scripts/ab_variants.py drude_chains > $O/swizzled_$i.jsonl 2>> $O/err.txt; echo "swz $i rc=$?" >> $O/status.txt
cp scripts/exp/libconstv_noswz.so openmm_constv_b200/libconstv_b200.so
AB_CHAINS=1,6 timeout 300 python scripts/ab_variants.py drude_chains > $O/linear_$i.jsonl 2>> $O/err.txt; echo "lin $i rc=$?" >> $O/status.txt
done
cp /tmp/lib_swz.so openmm_constv_b200/libconstv_b200.so
echo done >> $O/status.txt
