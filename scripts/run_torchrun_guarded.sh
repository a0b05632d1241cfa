#!/bin/bash
# usage: run_torchrun_guarded.sh N LIMIT_SECONDS bench-args...  — the driver's own launch line
# (python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ...) in its own process group,
# which is killed as a whole if it exceeds the limit (a hung collective must not eat the GPU budget).
N=$1; LIMIT=$2; shift 2
setsid python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29544} bench.py --gpus $N "$@" > gpurun_out/torchrun_$N.log 2>&1 &
PID=$!
( sleep $LIMIT; echo "watchdog: killing process group $PID"; kill -9 -- -$PID 2>/dev/null ) &
WD=$!
wait $PID; rc=$?
kill $WD 2>/dev/null
tail -${TAILN:-4} gpurun_out/torchrun_$N.log | cut -c1-900
exit $rc
