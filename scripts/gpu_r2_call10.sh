#!/bin/bash
# round 2, call 10: launch list of the x9 inference forward (128^2 and 216^2 tiles); stack-mode epilogue experiment bits
set -u
OUT=gpurun_out/r2c10
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for T in 128 216; do
  TB=8; [ $T = 216 ] && TB=3
  SR_TILE=$T SR_TILE_BATCH=$TB timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/infer_launches_$T.csv python scripts/profile_infer.py > $OUT/infer_$T.log 2>&1
  tail -2 $OUT/infer_$T.log | tee -a $OUT/summary.txt
  python scripts/summarize_launches.py $OUT/infer_launches_$T.csv 40 > $OUT/infer_summary_$T.txt 2>&1
  head -30 $OUT/infer_summary_$T.txt | tee -a $OUT/summary.txt
done
SR_DBG=0,1 SR_CTAS=0 SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/halo_trace.py > $OUT/halo_trace.txt 2>&1
