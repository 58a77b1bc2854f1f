"""Compose profiles/r02_k2b_epilogue.txt from the artefacts a GPU run leaves under gpurun_out/ (launch lists in CSV, timelines)."""
import csv, sys, os
G = "gpurun_out"

def search_launches(f, which):
    hdr = None; rows = []
    for r in csv.reader(open(f)):
        if 'Kernel Name' in r: hdr = r; continue
        if hdr and len(r) == len(hdr): rows.append(r)
    ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    seq = [(r[ki].split('(')[0].replace('void ', '').replace('sky::', '').replace('_kernel', '')[:22], float(r[vi].replace(',', '')) / 1000) for r in rows]
    idx = [i for i, (k, v) in enumerate(seq) if k.startswith('pack_queries')]
    last = seq[idx[which]:idx[which + 1]] if which + 1 < len(idx) else seq[idx[which]:]
    tot = {}
    for k, v in last: tot[k] = tot.get(k, 0) + v
    return ' | '.join(f"{k} {v:.0f}" for k, v in last), {k: round(v) for k, v in tot.items()}, round(sum(tot.values()))

print("""K2b (tc_batch.cu) on small shards: where the time of a search goes, before and after the last session of round 2.
Launch lists: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_batch|merge_phase|batch_|pack_"
              python scripts/time_search.py --n <rows> --q <Q> --k <k> --metric <m> --path batch --steps 1   (one complete search, device us;
              cold cache and serialised under ncu: compare shares, the bench lines are never taken under a profiler)
Timelines:    tools/trace_tb_phase.py <phase> (experiments build, SKY_TB_DEBUG=32): %clock of CTA 0's epilogue warps per visit
              (one visit = 128 bank rows x 256 queries = 48 MMAs of 128 x 256 x 16 = 6144 tensor-pipe cycles):
              start | wait for the accumulator | filter + survivors ("chunks") | write-back ("tail")
""")
for title, f, w in [("C3's shard at 8 GPUs (1.25 M x 768, Q = 4096, L2, k = 100): 4 epilogue warps + survivor queue", f'{G}/d_launches_1250000.csv', 2),
                    ("same, 8 epilogue warps", f'{G}/f8_launches_1.csv', 3),
                    ("same, 8 epilogue warps + phase merges with independent loads in flight (final)", f'{G}/h_launches_1.csv', 3),
                    ("experiment, dropped: the above + warp-aggregated (match.any) histogram atomics in the radix select -- the merges get slower", f'{G}/g_launches_1.csv', 3),
                    ("C4's share at 8 GPUs (12.5 M x 768, Q = 1000, cosine, k = 1000): 4 epilogue warps + survivor queue", f'{G}/d_launches_12500000.csv', 2),
                    ("same, 8 epilogue warps", f'{G}/f8_launches_2.csv', 3),
                    ("same, final", f'{G}/h_launches_2.csv', 3),
                    ("experiment, dropped: match.any histogram atomics", f'{G}/g_launches_2.csv', 3)]:
    if not os.path.exists(f): continue
    line, tot, sm = search_launches(f, w)
    print("== " + title)
    print("   launch order (us): " + line)
    print(f"   totals: {tot}  sum {sm} us\n")
for title, f, rows in [("phase 1 (one tile per CTA, ~170 survivors per visit), 4 epilogue warps + queue", f'{G}/e_trace_c3g8_p1.txt', (2, 8)),
                       ("phase 1, 8 epilogue warps", f'{G}/f_trace_c3g8_p1.txt', (2, 8)),
                       ("last phase (45 tiles per CTA), 4 epilogue warps: the accumulator is waiting (acc+ ~300), a visit takes ~6.9 k cycles", f'{G}/e_trace_c3g8_p4.txt', (20, 26)),
                       ("last phase, 8 epilogue warps: the epilogue waits ~1 k cycles for the accumulator, a visit takes ~6.3 k cycles = its MMAs", f'{G}/f_trace_c3g8_p4.txt', (20, 26))]:
    if not os.path.exists(f): continue
    L = open(f).read().splitlines()
    print("== timeline, " + title)
    for l in L[rows[0]:rows[1]]: print("   " + l[:200])
    print()
