#!/bin/bash
# round 2, last session: final check as the driver runs it (full GPU suite, smoke, default bench, reference arm), then the
# bench lines / launch lists / one ncu --set full capture of the kernels this session changed (K2b, K5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 1100 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/g_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -4 gpurun_out/g_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/g_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/g_bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/g_bench_ref.err
for wl in c3g8 c4g8 c5; do
  timeout 400 python bench.py --workload $wl --also none --no-cpu > gpurun_out/g_bench_$wl.json 2> gpurun_out/g_bench_$wl.err; echo "bench $wl rc=$? t=$(( $(date +%s) - T0 ))"
done
python - <<'PY'
import json
def show(n, r):
    if not r or 'skipped' in r: print(n, r); return
    rf = r['roofline']
    print(f"{n:6s} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')} clocks={r['clocks']['sm_mhz']} {r['clocks']['reasons']}")
try:
    d = json.loads(open('gpurun_out/g_bench_default.json').read().strip().splitlines()[-1])
    show('c2', d)
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
    for n, r in d.get('also', {}).items(): show(n, r)
    print('ref', open('gpurun_out/g_bench_ref.json').read()[:200])
except Exception as e:
    print('summary failed', e)
for wl in ('c3g8', 'c4g8', 'c5'):
    try:
        show(wl, json.loads(open(f'gpurun_out/g_bench_{wl}.json').read().strip().splitlines()[-1]))
    except Exception as e:
        print(wl, 'failed', e)
PY
i=0
for wl in "--n 1250000 --q 4096 --k 100 --metric MSE" "--n 12500000 --q 1000 --k 1000 --metric cosine"; do
  i=$((i+1))
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_batch|merge_phase|batch_|pack_" -c 100 --csv --log-file gpurun_out/g_launches_$i.csv python scripts/time_search.py $wl --path batch --steps 1 > gpurun_out/g_ncu_list_$i.log 2>&1; echo "ncu list $i rc=$? t=$(( $(date +%s) - T0 ))"
done
# the last (largest) phase of the last search of C3's 8-GPU shard: 4 searches x 5 scoring launches -> skip 19
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_batch -s 19 -c 1 -f -o gpurun_out/g_ncu_c3g8_tc_batch python scripts/time_search.py --n 1250000 --q 4096 --k 100 --metric MSE --path batch --steps 1 > gpurun_out/g_ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s) - T0 ))"
