"""Summarise an `ncu --page source --csv` export: executed warp-instructions by opcode
and stall-sample totals, for one kernel.  Usage: python tools/ncu_mix.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ops, samples, stalls = Counter(), Counter(), Counter()
total = 0
seen_sass = False
for r in rows[hdr_i + 1:]:
    if len(r) <= iE or not r[0].strip():
        continue
    src = r[iS].strip()
    try:
        e = int(float(r[iE] or 0)); n = int(float(r[iN] or 0))
    except ValueError:
        continue
    if not src or not src[0].isalpha() and not src.startswith("@"):
        continue
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG"):
        base = ".".join(op.split(".")[:1] + [p for p in op.split(".")[1:] if p in ("64", "128")])
    ops[base] += e; samples[base] += n; total += e
    for i in stall_cols:
        try:
            stalls[h[i]] += int(float(r[i] or 0))
        except ValueError:
            pass
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"total warp-instructions {total}")
for k, v in ops.most_common(top):
    print(f"{k:14s} {v:12d} {100.0 * v / total:6.2f}%   samples {samples[k]}")
ts = sum(stalls.values())
print("stall samples:", ", ".join(f"{k[6:]} {100.0 * v / ts:.1f}%" for k, v in stalls.most_common(8)))
