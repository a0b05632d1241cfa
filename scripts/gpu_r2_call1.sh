#!/bin/bash
# round 2, call 1: full GPU suite incl. the new full-size parity tests, default bench (all sub-lines), reference arm
set -u
OUT=gpurun_out/r2c1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_fullsize_parity.py > $OUT/pytest_gpu.log 2>&1
echo "pytest(old suite) exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q --timeout 900 -s > $OUT/pytest_fullsize.log 2>&1
echo "pytest(fullsize) exit $?" | tee -a $OUT/summary.txt
tail -30 $OUT/pytest_fullsize.log | tee -a $OUT/summary.txt
cp gpurun_out/parity_*.txt $OUT/ 2>/dev/null
timeout -s KILL 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout -s KILL 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/bench.err | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2c1/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("other mode", d.get("other_launch_mode"))
r = d["roofline"]; print("roofline", r["kernel"], r["achieved"], r["frac"], r["step_frac_of_tensor_roofline"]); print("hbm", r.get("memory_bound_families"))
print("g_forward", d.get("g_forward")); print("comparator", d.get("gpu_comparator"))
print("edsr", d["edsr"]); print("inference", d["inference"]); print("cpu", d["cpu_baseline"])
PY
