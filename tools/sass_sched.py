"""Static schedule of a kernel from `cuobjdump -sass`: per basic region, instruction counts by class and
the sum of the compile-time stall counts (the single-warp issue floor of the region).
Usage: python tools/sass_sched.py lib.so <mangled-substring> [lo hi]   (lo/hi: hex address window)"""
import re, subprocess, sys
from collections import Counter
so, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, ins = None, []
lines = txt.splitlines()
i = 0
while i < len(lines):
    l = lines[i]
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); i += 1; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", l)
    if m and cur and pat in cur:
        addr, text, w0 = int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)
        m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        w1 = int(m2.group(1), 16)
        stall = (w1 >> 41) & 0xF; yld = (w1 >> 45) & 1; wbar = (w1 >> 46) & 7; rbar = (w1 >> 49) & 7; wait = (w1 >> 52) & 0x3F
        ins.append((addr, text, stall, yld, wbar, rbar, wait))
        i += 2; continue
    i += 1
sel = [x for x in ins if lo <= x[0] < hi]
ops = Counter(); st = Counter(); tot = 0
for addr, text, stall, yld, wbar, rbar, wait in sel:
    t = text.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += 1; st[op] += stall; tot += stall
    if "-v" in sys.argv:
        print(f"{addr:06x} s{stall:2d} y{yld} w{wbar} r{rbar} m{wait:02x}  {text}")
print(f"{len(sel)} instructions, sum of stall counts {tot}")
for k, v in ops.most_common(20):
    print(f"  {k:10s} {v:5d}  stall {st[k]:5d}  avg {st[k]/v:.2f}")
