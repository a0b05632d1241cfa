#!/bin/bash
set -u
OUT=gpurun_out/call7
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -6 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench.json 2> $OUT/bench.err
SR_VGG_ASYNC=0 timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench_novgg.json 2> $OUT/bench_novgg.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_novgg.json"):
    try:
        d = json.loads(open("gpurun_out/call7/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/bench.err
SR_PROFILE=1 SR_ONLY=nothing timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"thin_cin|thin_wgrad_ci|act_bwd_ps2" -c 6 --profile-from-start off -o $OUT/thin_full python scripts/profile_step.py > $OUT/ncu_thin.log 2>&1
ncu -i $OUT/thin_full.ncu-rep --page raw --csv > $OUT/thin_full_raw.csv 2>/dev/null
tail -3 $OUT/ncu_thin.log | tee -a $OUT/summary.txt
