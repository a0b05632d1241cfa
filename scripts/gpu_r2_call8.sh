#!/bin/bash
# round 2, call 8: halo conv — row-coalesced (quad-transposed) bf16 stores, stack mode (N = 192) for 64 output channels
set -u
OUT=gpurun_out/r2c8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_fused_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_kernels.log 2>&1
echo "pytest(kernels) exit $?" | tee $OUT/summary.txt
tail -15 $OUT/pytest_kernels.log | tee -a $OUT/summary.txt
SR_ONLY=G.K timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1; cat $OUT/conv_bench.txt | tee -a $OUT/summary.txt
SR_HALO_STACK=0 SR_ONLY=G.K timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench_nostack.txt 2>&1; cat $OUT/conv_bench_nostack.txt | tee -a $OUT/summary.txt
SR_DBG=0 SR_CTAS=0,100 SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/halo_trace.py > $OUT/halo_trace.txt 2>&1
echo "trace exit $?" | tee -a $OUT/summary.txt
