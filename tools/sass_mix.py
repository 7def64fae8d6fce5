"""Aggregate an `ncu --page source --csv` export by SASS opcode.
usage: python tools/sass_mix.py file.csv [units]   (units = evals the launch processed)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[hi + 1:]:
    if len(r) <= iN or not r[iN].isdigit():
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    op = ".".join(src.split()[0].split(".")[:3]) if src else "?"
    n, s = int(r[iN]), int(r[iSamp] or 0)
    ops[op] += n
    samp[op] += s
    tot += n
    tots += s
print(f"total warp-inst {tot}  per unit {tot / units:.1f}  samples {tots}")
for op, n in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(f"{op:26s} {n / tot * 100:6.2f}%  per unit {n / units:7.1f}   stall samples {samp[op] / max(tots, 1) * 100:6.2f}%")
