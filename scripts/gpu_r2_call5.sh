#!/bin/bash
# round 2, call 3: band-path local-attention chain + conv epilogue pooling partials
set -u
OUT=gpurun_out/r2c5
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_fused_kernels.py tests/test_gpu_loss_cgam_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_kernels.log 2>&1
echo "pytest(kernels) exit $?" | tee $OUT/summary.txt
tail -30 $OUT/pytest_kernels.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py > $OUT/la_bench.txt 2>&1; cat $OUT/la_bench.txt | tee -a $OUT/summary.txt
SR_LA_BAND=0 timeout -s KILL 300 python scripts/la_bench.py > $OUT/la_bench_tile.txt 2>&1; head -2 $OUT/la_bench_tile.txt | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/diag_grad_err.py > $OUT/diag_grad.txt 2>&1; cat $OUT/diag_grad.txt | tee -a $OUT/summary.txt
SR_LA_BAND=0 timeout -s KILL 300 python scripts/diag_grad_err.py > $OUT/diag_grad_tile.txt 2>&1; cat $OUT/diag_grad_tile.txt | tee -a $OUT/summary.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_fullsize_parity.py --deselect tests/test_gpu_loss_cgam_kernels.py --deselect tests/test_gpu_fused_kernels.py > $OUT/pytest_gpu.log 2>&1
echo "pytest(rest) exit $?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q --timeout 900 -s > $OUT/pytest_fullsize.log 2>&1
echo "pytest(fullsize) exit $?" | tee -a $OUT/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|AssertionError:" $OUT/pytest_fullsize.log | tee -a $OUT/summary.txt
cp gpurun_out/parity_*.txt $OUT/ 2>/dev/null
timeout -s KILL 900 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 16 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/bench.err | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2c5/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("other mode", d.get("other_launch_mode"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  %-22s %4d launches %7.3f ms" % (k, v["launches_per_step"], v["ms_per_step"]), {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs") and b})
PY
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 70 > $OUT/launches_summary.txt 2>&1
head -75 $OUT/launches_summary.txt | tee -a $OUT/summary.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:la_.*band -s 60 -c 2 -o $OUT/ncu_la_band python scripts/la_bench.py > $OUT/ncu_la_band.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:la_bwd_apply_mma -s 30 -c 1 -o $OUT/ncu_la_apply python scripts/la_bench.py > $OUT/ncu_la_apply.log 2>&1
ls -la $OUT/*.ncu-rep | tee -a $OUT/summary.txt
