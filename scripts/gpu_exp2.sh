#!/bin/bash
set -u
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('   value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3))
"; }
run SKY_TC_DEBUG=3 SKY_TC_GRID=74
run SKY_TC_DEBUG=3 SKY_TC_GRID=37
run SKY_TC_DEBUG=3 SKY_TC_GRID=128
run SKY_TC_DEBUG=7
run SKY_TC_DEBUG=7 SKY_TC_GRID=128
run SKY_TC_DEBUG=7 SKY_TC_STAGES=3
