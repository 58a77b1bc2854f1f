"""Per-kernel counts of the SASS opcodes that prove which hardware paths the shipped library uses
(tcgen05 = UTCHMMA / UTCBAR / LDTM / STTM, TMA = UTMALDG, bulk copy = UBLKCP, cp.async = LDGSTS, cluster = UCGABAR).

    python tools/sass_opcodes.py [path/to/libskysearch.so] > profiles/rNN_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "sky_embeddings_b200", "libskysearch.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "UCGABAR", "SYNCS", "HMMA", "FFMA", "HFMA2", "ATOMS", "REDG", "MEMBAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        if op in OPS:
            counts[kern][op + (".2CTA" if ".2CTA" in m.group(2) else "")] += 1
print(f"# {os.path.basename(lib)}: cubins for {', '.join(arch)}; per kernel: SASS instructions, then counts of the opcodes of interest")
for k, c in counts.items():
    print(f"{k}: {total[k]} instr; " + (", ".join(f"{o} {n}" for o, n in sorted(c.items())) or "-"))
