#!/bin/bash
# round 2: targeted GPU tests after a change + a few bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_properties.py tests/test_gpu_cli.py -m gpu -q -k "batch or fullsize or tensor_paths or merge or cli_tile" --timeout 400 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_check.log | cut -c1-200
for wl in c3g8 c3; do
timeout 400 python bench.py --workload $wl --also none --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
python - $wl <<'PY'
import json, sys
wl = sys.argv[1]
try:
    r = json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"{wl} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} clk={r['clocks']['sm_mhz']}")
except Exception as e:
    print('summary failed', e)
PY
done
