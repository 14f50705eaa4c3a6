#!/bin/bash
O=gpurun_out/s9; mkdir -p $O
timeout 900 python scripts/ab_variants.py settle > $O/ab_settle.jsonl 2> $O/ab_settle.err; echo "ab_settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py settle 4000000 > $O/ab_settle_4m.jsonl 2>> $O/ab_settle.err; echo "ab_settle4m rc=$?" >> $O/status.txt
echo done >> $O/status.txt
