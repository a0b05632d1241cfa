#!/bin/bash
# round 2, call 9: store-pattern probe; full GPU suite + bench with stack mode / quad stores
set -u
OUT=gpurun_out/r2c9
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SR_LIB_PATH=build/probes/libsradsgan_b200.so timeout -s KILL 300 python scripts/store_probe.py > $OUT/store_probe.txt 2>&1
cat $OUT/store_probe.txt | tee $OUT/summary.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 16 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2c9/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  %-22s %4d launches %7.3f ms" % (k, v["launches_per_step"], v["ms_per_step"]), {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs") and b})
PY
