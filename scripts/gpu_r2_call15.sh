#!/bin/bash
# round 2, call 15: programmatic dependent launch on the conv / local-attention kernels: tests, A/B bench
set -u
OUT=gpurun_out/r2c15
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -6 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
for PDL in 1 0 1 0; do
  SR_PDL=$PDL timeout -s KILL 600 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 24 > $OUT/bench_pdl$PDL.json 2> $OUT/bench.err
  python - <<PY | tee -a $OUT/summary.txt
import json
d = json.loads(open("$OUT/bench_pdl$PDL.json").read().strip().splitlines()[-1])
print("SR_PDL=$PDL ms/step %.3f img/s %.1f e2e %.1f launches %d eager %.2f ms" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["other_launch_mode"]["ms_per_step"]))
PY
done
SR_TILE=128 SR_TILE_BATCH=8 timeout -s KILL 300 python scripts/profile_infer.py 2>&1 | tail -2 | tee -a $OUT/summary.txt
SR_PDL=0 SR_TILE=128 SR_TILE_BATCH=8 timeout -s KILL 300 python scripts/profile_infer.py 2>&1 | tail -2 | tee -a $OUT/summary.txt
