#!/bin/bash
# round 2, call A: full GPU test suite + the default bench line + the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import oracle.ref_harness as H; print('ref copy verified:', H.verify_ref())" > gpurun_out/ref.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    def show(n, r):
        if not r or 'skipped' in r: print(n, r); return
        rf = r['roofline']
        print(f"{n:6s} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac',0):.3f} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')} clocks={r['clocks']['sm_mhz']} {r['clocks']['reasons']}")
    show('c2', d)
    print('cpu', d.get('cpu_baseline'))
    for n, r in d.get('also', {}).items(): show(n, r)
    print('ref', open('gpurun_out/bench_ref.json').read()[:600])
except Exception as e:
    print('summary failed', e)
PY
