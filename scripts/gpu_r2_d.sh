#!/bin/bash
# round 2, session 3: pixel kernel 16 x 4 shape (four queries per pass), wider warp merge of K2b, launch lists of the K2b shards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "pixel or batch_path or config5 or config3" --timeout 250 --timeout-method=thread -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/d_pytest.log | cut -c1-200
timeout 300 python bench.py --workload c5 --also none --no-cpu --steps 20 --warmup 3 > gpurun_out/d_c5.json 2> gpurun_out/d_c5.err; echo "c5 rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/d_c5.err | cut -c1-300
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/d_c5.json').read().strip().splitlines()[-1])
    rf = r['roofline']; print('c5', 'ms/step', round(r['ms_per_step'], 3), 'kernel_ms', round(rf['kernel_ms'], 3), 'frac', round(rf['frac'], 3), 'value', round(r['value'], 1), r['clocks'])
except Exception as e:
    print('c5 summary failed', e)
PY
timeout 200 python scripts/ab_knobs.py c3g8r c4g8r 2>&1 | cut -c1-12,66-260
for wl in "--n 12500000 --q 1000 --k 1000 --metric cosine" "--n 1250000 --q 4096 --k 100 --metric MSE"; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_batch|merge_phase|batch_|pack_" -c 80 --csv --log-file gpurun_out/d_launches_$(echo $wl | cut -d' ' -f2).csv python scripts/time_search.py $wl --path batch --steps 1 > gpurun_out/d_ncu.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))"
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/d_launches_*.csv')):
    hdr=None; rows=[]
    for r in csv.reader(open(f)):
        if 'Kernel Name' in r: hdr=r; continue
        if hdr and len(r)==len(hdr): rows.append(r)
    if not hdr: print(f, 'no data'); continue
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    seq=[(r[ki].split('(')[0].replace('void ','').replace('sky::','')[:28], float(r[vi].replace(',',''))/1000) for r in rows]
    # last search = from the last pack_queries on
    idx=[i for i,(k,v) in enumerate(seq) if k.startswith('pack_queries')]
    last=seq[idx[-1]:] if idx else seq
    print(f, 'launches of the last search:', ' | '.join(f"{k} {v:.0f}" for k,v in last), ' total us', round(sum(v for k,v in last)))
PY
