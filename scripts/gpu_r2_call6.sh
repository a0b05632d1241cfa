#!/bin/bash
# round 2, call 6: halo conv role timeline (probes build) + check of the vectorised 1x1-weight staging of the la_* kernels
set -u
OUT=gpurun_out/r2c6
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/halo_trace.py > $OUT/halo_trace.txt 2>&1
echo "trace exit $?" | tee $OUT/summary.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_fused_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_kernels.log 2>&1
echo "pytest(kernels) exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_kernels.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py > $OUT/la_bench.txt 2>&1; cat $OUT/la_bench.txt | tee -a $OUT/summary.txt
SR_ONLY=G.K timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1; cat $OUT/conv_bench.txt | tee -a $OUT/summary.txt
