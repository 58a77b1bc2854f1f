#!/bin/bash
set -u
mkdir -p gpurun_out
ARGS="--workload ${WL:-c3s} --steps 2 --warmup 3 --no-cpu ${EXTRA:-}"
timeout 600 python bench.py $ARGS > gpurun_out/plain_list.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tc_batch|merge_|pack_queries|batch_|init_state|stream_search|tc_search|pixel" -c 600 --csv --log-file gpurun_out/launches_${WL:-c3s}.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_list.log
