#!/bin/bash
# One gpurun --gpus N call: torchrun bench at N ranks, plus single-GPU SIMT-regime bench lines.
set -u
mkdir -p gpurun_out
N=${NGPU:-2}
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
for wl in q1 q4; do for dt in bf16 fp32; do
  echo "== simt $wl $dt"
  timeout 300 python bench.py --workload $wl --bank-dtype $dt --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${wl}_${dt}.json 2> gpurun_out/bench_${wl}_${dt}.err
  echo "rc=$?"; cat gpurun_out/bench_${wl}_${dt}.json; tail -3 gpurun_out/bench_${wl}_${dt}.err
done; done
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "rc=$?"; cat gpurun_out/bench_ref.json
