#!/bin/bash
# round 2, call 12: discriminator attention primitives (csrc/cbam.cu)
set -u
OUT=gpurun_out/r2c12
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_cbam_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_cbam.log 2>&1
echo "pytest(cbam) exit $?" | tee $OUT/summary.txt
tail -25 $OUT/pytest_cbam.log | tee -a $OUT/summary.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_gpu_cbam_kernels.py > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 16 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2c12/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("other mode", d.get("other_launch_mode"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  %-22s %4d launches %7.3f ms" % (k, v["launches_per_step"], v["ms_per_step"]), {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs") and b})
PY
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 80 > $OUT/launches_summary.txt 2>&1
head -85 $OUT/launches_summary.txt | tee -a $OUT/summary.txt
