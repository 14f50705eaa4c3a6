#!/bin/bash
O=gpurun_out/s16; mkdir -p $O
timeout 600 python scripts/ab_variants.py chains > $O/ab_chains.jsonl 2> $O/ab.err; echo "ab rc=$?" >> $O/status.txt
echo done >> $O/status.txt
