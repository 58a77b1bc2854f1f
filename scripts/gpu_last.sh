#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "adversarial_order_exercises or golden_simsearch or multi_query or nan_rows or large_k or mae_simsearch_mirror" 2>&1 | tail -3
for cfg in "q1 fp32 --weighted" "q1 bf16" "l64 fp32 --weighted"; do
set -- $cfg
timeout 200 python bench.py --workload $1 --bank-dtype $2 ${3:-} --steps 40 --warmup 5 --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw.json")); r=d["roofline"]
    print("$cfg kernel_ms=%.4f GB/s=%.0f frac=%.3f step_ms=%.4f" % (r["kernel_ms"], r["achieved"], r["frac"], d["ms_per_step"]))
except Exception as e: print("ERR", e, open("gpurun_out/sw.err").read()[-200:])
PY
done
