#!/bin/bash
set -u
mkdir -p gpurun_out
ARGS="--workload ${WL:-q1} --bank-dtype ${DT:-bf16} --path simt --steps 10 --warmup 3 --no-cpu"
timeout 300 python bench.py $ARGS > gpurun_out/plain_stream.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:stream_search" -s 3 -c 1 -f -o gpurun_out/prof_stream_${WL:-q1}_${DT:-bf16} python bench.py $ARGS > gpurun_out/ncu_stream.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_stream.log
