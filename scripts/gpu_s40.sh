#!/bin/bash
O=gpurun_out/s40; mkdir -p $O
for m in settle_minbd drude_minbd; do
  timeout 900 python scripts/ab_variants.py $m 1000000 mixed > $O/ab_${m}_mixed.jsonl 2> $O/ab_${m}_mixed.err; echo "$m rc=$?" >> $O/status.txt
done
echo done >> $O/status.txt
