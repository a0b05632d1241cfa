#!/bin/bash
# round 2, call 23: programmatic dependent launch really switched on (SR_PDL is now forwarded): step A/B + GPU tests under it
set -u
OUT=gpurun_out/r2c23
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for PDL in 1 0; do
  SR_PDL=$PDL timeout -s KILL 300 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 24 > $OUT/bench_pdl$PDL.json 2> $OUT/bench.err
  python - <<PY | tee -a $OUT/summary.txt
import json
d = json.loads(open("$OUT/bench_pdl$PDL.json").read().strip().splitlines()[-1])
print("SR_PDL=$PDL ms/step %.3f img/s %.1f e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
done
SR_PDL=1 timeout -s KILL 400 python -m pytest tests/test_gpu_fused_kernels.py tests/test_gpu_conv_kernels.py tests/test_gpu_model_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q --timeout 300 -x > $OUT/pytest_pdl.log 2>&1
echo "pytest(SR_PDL=1) exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_pdl.log | tee -a $OUT/summary.txt
