#!/bin/bash
set -u
OUT=gpurun_out/call10
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/call10/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
PY
tail -3 $OUT/bench.err
