"""Join an `ncu --page source --csv` (SASS) export with `nvdisasm -gi` line info and print
the executed warp-instructions and stall samples per CUDA source line.
usage: python tools/line_profile.py ncu_sass.csv all.sass <mangled-kernel-substring> [units] [top]"""
import collections
import csv
import re
import sys

ncu_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

line_of = {}
cur = None
inside = False
fresh = True
for ln in open(sass):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kname in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:                      # innermost frame comes first
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            fresh = False
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
        fresh = True

rows = list(csv.reader(open(ncu_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iA, iN, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[iA], 16), int(r[iN]), int(r[iS] or 0)) for r in rows[hi + 1:]
        if len(r) > iN and r[iN].isdigit()]
base = data[0][0]
agg_n, agg_s = collections.Counter(), collections.Counter()
tot = tots = 0
for a, n, s in data:
    key = line_of.get(a - base, ("?", 0))
    agg_n[key] += n
    agg_s[key] += s
    tot += n
    tots += s
src_cache = {}
def src(key):
    f, l = key
    import glob, os
    if f not in src_cache:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "**", f),
                      recursive=True)
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    L = src_cache[f]
    return L[l - 1].strip()[:90] if 0 < l <= len(L) else ""
print(f"total warp-inst {tot} ({tot / units:.1f} per unit), samples {tots}")
for key, n in agg_n.most_common(top):
    print(f"{key[0]:>16s}:{key[1]:<4d} {n / units:7.1f}/unit {n / tot * 100:5.1f}%  stall {agg_s[key] / max(tots, 1) * 100:5.1f}%  | {src(key)}")
