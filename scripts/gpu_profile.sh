#!/bin/bash
# One gpurun call: bench (plain), then the ncu launch list and one --set full capture of the scoring kernel.
set -u
mkdir -p gpurun_out
ARGS="--steps ${STEPS:-20} --warmup 3 --no-cpu ${BENCH_ARGS:-}"
KREGEX='regex:tc_search|simt_search|merge_|pack_queries|init_state'
echo "== bench (full, with cpu baseline)"; timeout 900 python bench.py --steps 200 --warmup 10 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== plain run for ncu"; timeout 600 python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/ncu_list.log
timeout 600 python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${PROF_KERNEL:-tc_search}" -s 5 -c 2 -f -o gpurun_out/prof_${PROF_NAME:-tc} python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
