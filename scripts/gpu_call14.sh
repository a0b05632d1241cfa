#!/bin/bash
set -u
OUT=gpurun_out/call14
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_fused_kernels.py -q --timeout 120 > $OUT/pytest_fused.log 2>&1
echo "pytest fused exit $?" | tee $OUT/summary.txt
SR_LA_GATE_TAIL=0 timeout -s KILL 300 python -m pytest tests/test_gpu_fused_kernels.py -q --timeout 120 -k la_chain > $OUT/pytest_fused_notail.log 2>&1
echo "pytest fused (no tail) exit $?" | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
for rep in 1 2; do
for cfg in "SR_LA_GATE_TAIL=1" "SR_LA_GATE_TAIL=0" "SR_LA_GATE_TAIL=1 SR_LA_SLICE_PX=256" "SR_LA_GATE_TAIL=1 SR_LA_SIDE=0" "SR_LA_GATE_TAIL=1 SR_WGRAD_ASYNC=0"; do
  env $cfg timeout -s KILL 600 python bench.py --no-cpu-baseline --no-inference --no-edsr --steps 16 > $OUT/b.json 2> $OUT/b.err
  python - "$cfg" <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/call14/b.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "img/s %.1f" % d["value"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
done
