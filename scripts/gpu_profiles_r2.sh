#!/bin/bash
# Round-2 profile set: launch list of the default C2 step, ncu --set full captures of the kernels VERDICT named
# (K2w2 weighted pair kernel, K2, K5 at Q = 1, K1 with weights on a bf16 bank) and the launch list of a small-shard
# K2b search.  Every ncu run follows a plain run of the same command that exited 0.
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/prof
O=gpurun_out/prof
A="--also none --steps 4 --warmup 3 --no-cpu"
list() { # name, kernel regex, count, args...
  name=$1; kre=$2; cnt=$3; shift 3
  timeout 600 python bench.py "$@" > $O/plain_list_$name.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$kre" -c $cnt --csv --log-file $O/launches_$name.csv python bench.py "$@" > $O/ncu_list_$name.log 2>&1
  echo "list $name rc=$?"
}
prof() { # name, kernel regex, skip, args...
  name=$1; kre=$2; skip=$3; shift 3
  timeout 600 python bench.py "$@" > $O/plain_$name.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c 1 -f -o $O/ncu_$name python bench.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
list c2 "tc_search|merge_|pack_queries|init_state" 120 $A
list c2w "tc_weighted|merge_|pack_weighted|init_state" 120 --workload c2w $A
list c3g8 "tc_batch|merge_phase|batch_|pack_queries" 400 --workload c3g8 --also none --steps 2 --warmup 3 --no-cpu
prof c2w_tc_weighted2 tc_weighted2 3 --workload c2w $A
prof c2_tc_search tc_search 5 $A
prof q1wb_stream stream_search 3 --workload q1wb $A
prof c5q1_pixel pixel_search 2 --workload c5q1s --also none --steps 2 --warmup 3 --no-cpu
for wl in q1wb q1 q4 c2mse l64 c3g8 c4g8 c5 c5q1; do
  timeout 900 python bench.py --workload $wl --also none --no-cpu > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"
done
ls -la $O | head -60
