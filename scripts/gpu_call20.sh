#!/bin/bash
set -u
OUT=gpurun_out/call20
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
for rep in 1 2 3; do
for cfg in "SR_WGRAD_STREAMS=2" "SR_WGRAD_STREAMS=1"; do
  env $cfg timeout -s KILL 600 python bench.py --no-cpu-baseline --no-inference --no-edsr --steps 16 > $OUT/b.json 2> $OUT/b.err
  python - "$cfg" <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/call20/b.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "img/s %.1f" % d["value"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
done
