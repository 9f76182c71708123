#!/usr/bin/env python3
"""Aggregate an ncu SASS source page (`ncu -i X.ncu-rep --page source --csv --print-source sass`) per CUDA source
line, using `nvdisasm -g -c <cubin>` output for the address -> line map.
usage: sass_by_line.py sass.csv disasm.txt kernel_substring [launch_index] [top_n, 0 = all lines]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, kname = sys.argv[1:4]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 60

# 1. ordered list of source lines per instruction of the kernel from nvdisasm
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if ".text." in l and kname in l and l.startswith("//-----"))
cur = None
inline = None
instr_lines = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        instr_lines.append(cur)

rows = list(csv.reader(open(sass_csv)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[heads[launch]]
end = heads[launch + 1] - 1 if launch + 1 < len(heads) else len(rows)
body = [r for r in rows[heads[launch] + 1:end] if r and r[0].startswith("0x")]
ix = {n: h.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "Source")}
assert len(body) == len(instr_lines), (len(body), len(instr_lines))
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, ln in zip(body, instr_lines):
    v = [int(r[ix["Instructions Executed"]] or 0), int(r[ix["Thread Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)]
    for k in range(3):
        agg[ln][k] += v[k]
        tot[k] += v[k]
print(f"total warp-inst {tot[0]}  thread-inst {tot[1]}  samples {tot[2]}")
print(f"{'file:line':28s} {'warp-inst':>10s} {'%':>6s} {'thr/inst':>8s} {'samples':>8s} {'%':>6s}")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:(top_n if top_n > 0 else None)]:
    print(f"{str(ln[0]) + ':' + str(ln[1]):28s} {v[0]:10d} {100 * v[0] / tot[0]:6.2f} {v[1] / max(v[0], 1):8.1f} {v[2]:8d} {100 * v[2] / max(tot[2], 1):6.2f}")
