#!/bin/bash
set -u
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  for wl in ${WLS:-q1 q4}; do for dt in ${DTS:-bf16 fp32}; do
    env "$@" timeout 200 python bench.py --workload $wl --bank-dtype $dt --path simt --steps 30 --warmup 5 --no-cpu ${EXTRA:-} > gpurun_out/sw.json 2> gpurun_out/sw.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw.json")); r=d["roofline"]
    print("$label $wl $dt kernel_ms=%.4f GB/s=%.0f frac=%.3f step_ms=%.4f" % (r["kernel_ms"], r["achieved"], r["frac"], d["ms_per_step"]))
except Exception as e: print("$label $wl $dt ERR", e, open("gpurun_out/sw.err").read()[-300:])
PY
  done; done
}
run ${LABEL:-base} X=1
EXTRA=--weighted run ${LABEL:-base}_weighted X=1
