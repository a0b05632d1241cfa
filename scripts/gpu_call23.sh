#!/bin/bash
set -u
OUT=gpurun_out/call23
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_conv_kernels.py -q --timeout 60 -x > $OUT/pytest_conv.log 2>&1
echo "pytest conv exit $?" | tee $OUT/summary.txt
tail -4 $OUT/pytest_conv.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
SR_ONLY=G.K timeout -s KILL 300 python scripts/conv_bench.py 2>&1 | grep "G.K" | tee -a $OUT/summary.txt
SR_HALO_SPLITK=0 SR_ONLY=G.K timeout -s KILL 300 python scripts/conv_bench.py 2>&1 | grep "G.K" | tee -a $OUT/summary.txt
for rep in 1 2; do
for cfg in "SR_HALO_SPLITK=1" "SR_HALO_SPLITK=0"; do
  env $cfg timeout -s KILL 600 python bench.py --no-cpu-baseline --no-inference --no-edsr --steps 16 > $OUT/b.json 2> $OUT/b.err
  python - "$cfg" <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/call23/b.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "img/s %.1f" % d["value"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
done
