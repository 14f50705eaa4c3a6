#!/bin/bash
# round 2 session 69: smoke with the bulk-staged kernel check; the GPU suite once more on the last source
O=gpurun_out/s69; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
echo done >> $O/status.txt
