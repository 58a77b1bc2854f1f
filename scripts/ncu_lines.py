"""Aggregate `ncu --page source --csv` stall samples by CUDA source line, using nvdisasm -g line
markers of the kernel's cubin.  usage: ncu_lines.py <src.csv> <lib.so> <kernel-substring> [topn]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

csv_path, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, active = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m:
            active = kname in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur_line:
            line_of[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(csv_path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ns, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[hi + 1:] if len(r) > ns and r[0] != 'Address']
base = int(data[0][0], 16)
agg, inst = {}, {}
tot = 0
for r in data:
    off = int(r[0], 16) - base
    if off < 0:
        base = int(r[0], 16)     # second launch of the same kernel
        off = 0
    key = line_of.get(off, ("?", 0))
    agg[key] = agg.get(key, 0) + int(r[ns] or 0)
    inst[key] = inst.get(key, 0) + int(r[ie] or 0)
    tot += int(r[ns] or 0)
print("total samples", tot)
srcs = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    text = ""
    for d in ("sky_embeddings_b200/csrc",):
        p = os.path.join(d, f)
        if os.path.exists(p):
            srcs.setdefault(p, open(p).read().splitlines())
            if 0 < l <= len(srcs[p]):
                text = srcs[p][l - 1].strip()
    print(f"{v:7d} {100 * v / tot:5.1f}% inst={inst[(f, l)]:>10d} {f}:{l:<4d} {text[:100]}")
