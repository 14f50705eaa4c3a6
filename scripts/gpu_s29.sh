#!/bin/bash
O=gpurun_out/s29; mkdir -p $O
timeout 900 python scripts/ab_variants.py settle_chains > $O/ab_settle_chains.jsonl 2> $O/ab_settle_chains.err; echo "settle rc=$?" >> $O/status.txt
timeout 900 python scripts/ab_variants.py drude_chains > $O/ab_drude_chains.jsonl 2> $O/ab_drude_chains.err; echo "drude rc=$?" >> $O/status.txt
echo done >> $O/status.txt
