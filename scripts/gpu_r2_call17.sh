#!/bin/bash
# round 2, call 17: stack mode with 2-CTA clusters multicasting the weight stages
set -u
OUT=gpurun_out/r2c17
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -q --timeout 120 -x -k "halo" > $OUT/pytest_halo.log 2>&1
echo "pytest(halo) exit $?" | tee $OUT/summary.txt
tail -15 $OUT/pytest_halo.log | tee -a $OUT/summary.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_fused_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_kernels.log 2>&1
echo "pytest(kernels) exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_kernels.log | tee -a $OUT/summary.txt
SR_DBG=0 SR_CTAS=0,1,100 SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/halo_trace.py > $OUT/halo_trace.txt 2>&1
grep "lifetime" $OUT/halo_trace.txt | tee -a $OUT/summary.txt
for MODE in "SR_HALO_MCAST=1" "SR_HALO_MCAST=0" "SR_HALO_STACK_PAIRS=1"; do
  env $MODE SR_TILE=128 SR_TILE_BATCH=8 timeout -s KILL 300 python scripts/profile_infer.py 2>&1 | tail -2 | head -1 | tee -a $OUT/summary.txt
  env $MODE timeout -s KILL 600 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 24 > $OUT/bench_$MODE.json 2> $OUT/bench.err
  python - <<PY | tee -a $OUT/summary.txt
import json
d = json.loads(open("$OUT/bench_$MODE.json").read().strip().splitlines()[-1])
print("$MODE ms/step %.3f img/s %.1f e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
done
