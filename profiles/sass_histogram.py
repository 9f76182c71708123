#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libmarl_b200.so (cuobjdump -sass): the mnemonics that prove what a kernel runs on.
    python profiles/sass_histogram.py [path/to/libmarl_b200.so] > profiles/rNN_sass_opcode_histogram.txt
UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; UTMALDG / UTMASTG = TMA bulk tensor load / store; LDTM = tcgen05.ld;
UTCBAR = tcgen05.commit; SYNCS = mbarrier ops; FFMA2 = packed fp32 FMA; DFMA / DADD / DMUL = fp64."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dqn_marl_b200", "libmarl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "LDGSTS", "FFMA2", "FFMA", "DFMA", "DADD", "DMUL",
        "IMAD", "LOP3", "SHFL", "VOTE", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "BAR"]
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("mq::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        hist[kern][base] += 1
        hist[kern]["__total"] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            hist[kern]["UTCHMMA.2CTA"] += 1
cols = [k for k in KEYS if any(h[k] for h in hist.values())]
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{c[:8]:>8s}" for c in cols))
for k, h in hist.items():
    print(f"{k[:78]:78s} {h['__total']:7d} " + " ".join(f"{h[c]:8d}" if h[c] else f"{'.':>8s}" for c in cols))
