#!/bin/bash
# usage: run_ranks.sh N LIMIT_SECONDS script.py args...   — spawns N ranks of one node directly (env rendezvous on
# 127.0.0.1) and kills exactly those PIDs if they exceed the limit (a hung collective must not eat the GPU budget).
N=$1; LIMIT=$2; shift 2
export MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29533} WORLD_SIZE=$N
PIDS=()
for ((r=0; r<N; r++)); do
  RANK=$r LOCAL_RANK=$r python -W ignore "$@" > gpurun_out/rank$r.log 2>&1 &
  PIDS+=($!)
done
( sleep $LIMIT; echo "watchdog: killing ${PIDS[*]}"; kill -9 "${PIDS[@]}" 2>/dev/null ) &
WD=$!
rc=0
for p in "${PIDS[@]}"; do wait $p || rc=$?; done
kill $WD 2>/dev/null
for ((r=0; r<N; r++)); do echo "--- rank $r"; tail -${TAILN:-6} gpurun_out/rank$r.log | cut -c1-700; done
exit $rc
