#!/bin/bash
# One gpurun --gpus N call: torchrun bench lines at N ranks.
set -u
mkdir -p gpurun_out
N=${NGPU:-2}
run() { # name, args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/multi_${name}_n$N.json 2> gpurun_out/multi_${name}_n$N.err
  echo "$name N=$N rc=$?"; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version" gpurun_out/multi_${name}_n$N.err | tail -3
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/multi_${name}_n$N.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("  value=%.0f global_qps=%.0f e2e=%.0f step_ms=%.3f kernel_ms=%.3f %s frac=%.3f" % (d["value"], d["qps_global_bank"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["unit"], r["frac"]))
except Exception as e: print("ERR", e)
PY
}
run c2 --steps 200 --warmup 10 --no-cpu
run c4 --workload c4 --steps 5 --warmup 3 --no-cpu
run c3g8 --workload c3g8 --steps 10 --warmup 3 --no-cpu
