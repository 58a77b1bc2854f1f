#!/bin/bash
# full GPU test-suite + smoke + headline bench
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2.json")); r=d["roofline"]
print("c2 value=%.0f e2e=%.0f kernel_ms=%.4f frac=%.3f step_ms=%.4f cpu=%s" % (d["value"], d["e2e"]["value"], r["kernel_ms"], r["frac"], d["ms_per_step"], d.get("cpu_baseline",{}).get("value")))
PY
