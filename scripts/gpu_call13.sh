#!/bin/bash
set -u
OUT=gpurun_out/call13
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-edsr > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/call13/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"])
print("inference", d["inference"]["value"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  ", k, round(v["ms_per_step"], 3), v["launches_per_step"], round(v["tflops"] or 0, 1))
PY
tail -3 $OUT/bench.err
