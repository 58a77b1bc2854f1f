#!/bin/bash
# round 2 final check, as the driver will run it: full GPU suite, smoke, default bench, reference arm; then measurements
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/bench_ref.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    def show(n, r):
        if not r or 'skipped' in r: print(n, r); return
        rf = r['roofline']
        print(f"{n:6s} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')} clocks={r['clocks']['sm_mhz']} {r['clocks']['reasons']}")
    show('c2', d)
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
    for n, r in d.get('also', {}).items(): show(n, r)
    print('ref', open('gpurun_out/bench_ref.json').read()[:200])
except Exception as e:
    print('summary failed', e)
PY
