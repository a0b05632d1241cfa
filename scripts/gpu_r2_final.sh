#!/bin/bash
# round 2, final validation: full GPU suite, smoke, default bench (all sub-lines), per-shape conv bench, warm ncu of the RAB convs, launch list
set -u
OUT=gpurun_out/r2final
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
cp gpurun_out/parity_*.txt $OUT/ 2>/dev/null
timeout -s KILL 200 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
tail -2 $OUT/bench.err | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2final/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("other mode", d.get("other_launch_mode"))
r = d["roofline"]; print("roofline", r["kernel"], r["achieved"], r["frac"], r["step_frac_of_tensor_roofline"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  %-22s %4d launches %7.3f ms" % (k, v["launches_per_step"], v["ms_per_step"]), {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs") and b})
print("g_forward", d.get("g_forward")); print("comparator", d.get("gpu_comparator"))
print("edsr", d["edsr"]); print("inference", d["inference"]); print("cpu", d["cpu_baseline"])
PY
SR_ITERS=30 timeout -s KILL 200 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1; tail -1 $OUT/conv_bench.txt | tee -a $OUT/summary.txt
SR_ONLY=G.K timeout -s KILL 300 ncu --cache-control none --clock-control none -k regex:conv_halo_kernel --launch-skip 8 --launch-count 84 \
  --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file $OUT/ncu_halo_warm.csv python scripts/conv_bench.py > $OUT/ncu_halo_warm.log 2>&1
echo "ncu warm exit $?" | tee -a $OUT/summary.txt
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 80 > $OUT/launches_summary.txt 2>&1
head -12 $OUT/launches_summary.txt | tee -a $OUT/summary.txt
