#!/bin/bash
# tests + headline bench + streaming regime lines
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2.json")); r=d["roofline"]
print("c2 value=%.0f e2e=%.0f kernel_ms=%.4f frac=%.3f step_ms=%.4f" % (d["value"], d["e2e"]["value"], r["kernel_ms"], r["frac"], d["ms_per_step"]))
PY
WLS="q1 q4" bash scripts/gpu_sweep_stream.sh
