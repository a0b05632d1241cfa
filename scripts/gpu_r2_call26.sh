#!/bin/bash
# round 2, last call: the complete GPU suite + smoke on the final tree
set -u
OUT=gpurun_out/r2c26
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 400 -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -6 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
