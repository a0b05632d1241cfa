#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== gpu tests"; timeout -s KILL 900 python -W ignore -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
echo "== torch prof"; timeout -s KILL 400 python -W ignore scripts/torch_prof.py > gpurun_out/torch_prof.log 2>&1; grep "^==" gpurun_out/torch_prof.log
