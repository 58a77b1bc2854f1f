#!/bin/bash
# round 2: final check of HEAD as the driver runs it (full GPU suite, smoke, default bench, reference arm) + the K2b shard lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 330 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/g_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/g_pytest_gpu.log | cut -c1-200
timeout 100 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/g_smoke.log | cut -c1-200
timeout 200 python bench.py > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/g_bench_default.err
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - T0 ))"
for wl in c3g8 c4g8; do
  timeout 100 python bench.py --workload $wl --also none --no-cpu > gpurun_out/g_bench_$wl.json 2> gpurun_out/g_bench_$wl.err; echo "bench $wl rc=$? t=$(( $(date +%s) - T0 ))"
done
python - <<'PY'
import json
def show(n, r):
    if not r or 'skipped' in r: print(n, r); return
    rf = r['roofline']
    print(f"{n:6s} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} clocks={r['clocks']['sm_mhz']} {r['clocks']['reasons']}")
try:
    d = json.loads(open('gpurun_out/g_bench_default.json').read().strip().splitlines()[-1])
    show('c2', d)
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
    for n, r in d.get('also', {}).items(): show(n, r)
    print('ref', open('gpurun_out/g_bench_ref.json').read()[:160])
except Exception as e:
    print('summary failed', e)
for wl in ('c3g8', 'c4g8'):
    try: show(wl, json.loads(open(f'gpurun_out/g_bench_{wl}.json').read().strip().splitlines()[-1]))
    except Exception as e: print(wl, 'failed', e)
PY
