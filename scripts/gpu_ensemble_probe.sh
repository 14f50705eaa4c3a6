#!/bin/bash
# isolates the batched kernel's own overhead from the multi-slab memory layout
run() { python bench.py --steps 20000 --skip-cpu --skip-e2e --skip-ref-gpu "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', '->', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3))"; }
run --atoms 800000
run --ensemble 1 --atoms 800000
run --ensemble 2 --atoms 400000
run --ensemble 8 --atoms 100000
run --ensemble 8 --atoms 100000 --ctas-per-sm 5
run --ensemble 8 --atoms 100000 --ctas-per-sm 10
run --ensemble 8 --atoms 100000 --keep-in-l2
run --ensemble 64 --atoms 100000
