#!/bin/bash
# round 2: first run of the CTA-pair weighted kernel: diagnostic, its parity tests, the weighted C2 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/diag_k2w2.py > gpurun_out/diag_k2w2.log 2>&1; echo "diag rc=$?"; tail -12 gpurun_out/diag_k2w2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weighted_tensor or smoke or multi_query" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k2w2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_k2w2.log
timeout 300 python bench.py --workload c2w --also none --no-cpu --steps 20 --warmup 5 > gpurun_out/bench_c2w.json 2> gpurun_out/bench_c2w.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2w.err
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_c2w.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"c2w value={r['value']:.1f} ms={r['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} parity={r.get('parity')}")
except Exception as e:
    print('summary failed', e)
PY
