"""Print the hottest SASS lines of an `ncu --page source --csv` dump by warp-stall samples."""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
src, ns, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[hi + 1:] if len(r) > ns and r[0] != 'Address']
tot = sum(int(r[ns] or 0) for r in data)
print('total samples', tot, 'rows', len(data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
for r in sorted(data, key=lambda r: -int(r[ns] or 0))[:topn]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[ns]):7d} {100 * int(r[ns]) / tot:5.1f}% inst={r[ie]:>9s} {r[0][-5:]} {r[src][:80]:80s} {st}")
