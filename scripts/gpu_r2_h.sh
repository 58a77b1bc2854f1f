#!/bin/bash
# round 2, last session: phase merges with independent loads -- parity, then timing and launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_properties.py -q -x -k "batch or config3 or config4" --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/h_pytest.log | cut -c1-200
timeout 300 python scripts/ab_knobs.py c3g8r c4g8r c3r 2>&1 | cut -c1-12,66-260
i=0
for wl in "--n 1250000 --q 4096 --k 100 --metric MSE" "--n 12500000 --q 1000 --k 1000 --metric cosine"; do
  i=$((i+1))
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_batch|merge_phase|batch_|pack_" -c 100 --csv --log-file gpurun_out/h_launches_$i.csv python scripts/time_search.py $wl --path batch --steps 1 > gpurun_out/h_ncu_list_$i.log 2>&1; echo "ncu list $i rc=$? t=$(( $(date +%s) - T0 ))"
done
python - <<'PY'
import csv
for f in ('gpurun_out/h_launches_1.csv', 'gpurun_out/h_launches_2.csv'):
    hdr=None; rows=[]
    for r in csv.reader(open(f)):
        if 'Kernel Name' in r: hdr=r; continue
        if hdr and len(r)==len(hdr): rows.append(r)
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    seq=[(r[ki].split('(')[0].replace('void ','').replace('sky::','').replace('_kernel','')[:20], float(r[vi].replace(',',''))/1000) for r in rows]
    idx=[i for i,(k,v) in enumerate(seq) if k.startswith('pack_queries')]
    last=seq[idx[3]:] if len(idx) > 3 else seq[idx[-1]:]
    print(f, ' | '.join(f"{k} {v:.0f}" for k,v in last), ' sum', round(sum(v for k,v in last)))
PY
