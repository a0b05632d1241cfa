#!/bin/bash
# round 2, call 7: halo conv timeline with the epilogue experiment bits, launch gaps by global timer
set -u
OUT=gpurun_out/r2c7
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SR_DBG=0,1,2 SR_CTAS=0 SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/halo_trace.py > $OUT/halo_trace.txt 2>&1
echo "trace exit $?" | tee $OUT/summary.txt
