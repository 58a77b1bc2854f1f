#!/bin/bash
# Round profile set: one bench line per workload, then ncu launch list (C2) and --set full captures of the
# dominant kernel of each regime.  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out/prof
O=gpurun_out/prof
line() { # name, args...
  name=$1; shift
  timeout 900 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"
}
line c2 --steps 200 --warmup 10
line q1_fp32_weighted --workload q1 --bank-dtype fp32 --weighted --steps 50 --warmup 5 --no-cpu
line q1_bf16 --workload q1 --steps 50 --warmup 5 --no-cpu
line l64_fp32_weighted --workload l64 --bank-dtype fp32 --weighted --steps 50 --warmup 5 --no-cpu
line c2_weighted --weighted --steps 100 --warmup 5 --no-cpu
line c3 --workload c3 --steps 5 --warmup 3 --no-cpu
line c3_shard_of_8 --workload c3g8 --steps 5 --warmup 3 --no-cpu
line c4share --workload c4 --steps 5 --warmup 3 --no-cpu
line c5_q1 --workload c5q1 --steps 5 --warmup 3
line c5_q4 --workload c5 --steps 5 --warmup 3 --no-cpu
prof() { # name, kernel regex, skip, args...
  name=$1; kre=$2; skip=$3; shift 3
  timeout 600 python bench.py "$@" > $O/plain_$name.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c 1 -f -o $O/ncu_$name python bench.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
A="--steps 4 --warmup 3 --no-cpu"
timeout 600 python bench.py $A > $O/plain_list.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tc_search|stream_search|merge_|pack_queries|init_state" -c 200 --csv --log-file $O/launches_c2.csv python bench.py $A > $O/ncu_list.log 2>&1
echo "list rc=$?"
prof c2_tc_search tc_search 5 $A
prof q1_stream stream_search 3 --workload q1 --bank-dtype fp32 --weighted $A
prof c3_tc_batch tc_batch 26 --workload c3 --steps 1 --warmup 3 --no-cpu
prof c2w_tc_weighted tc_weighted 3 --weighted $A
prof c5_pixel pixel_search 2 --workload c5s --steps 2 --warmup 3 --no-cpu
ls -la $O
