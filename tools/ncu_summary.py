"""Summarise an .ncu-rep (read here, no GPU needed) into the text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1

writes <prefix>_metrics.txt (per-kernel raw metrics that matter for a streaming FP32 kernel:
duration, DRAM bytes, pipe and issue utilisation, occupancy limits, stall reasons) and
<prefix>_<kernel>_mix.txt (executed warp-instructions by opcode + stall-sample shares from the
source page; needs -lineinfo / --import-source on)."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep, prefix = sys.argv[1], sys.argv[2]
WANT = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|sm__warps_active\.avg\.pct|launch__registers_per_thread$|"
    r"launch__occupancy_limit_(registers|shared_mem|warps)|launch__shared_mem_per_block_dynamic|launch__grid_size|launch__block_size|"
    r"launch__waves_per_multiprocessor|sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_(alu|fma|lsu|xu)\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|lts__t_sector_hit_rate\.pct|sm__cycles_elapsed\.avg$|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio)")


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True, check=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
h, units = raw[0], raw[1]
with open(prefix + "_metrics.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none, report {rep}\n")
    for r in raw[2:]:
        name = r[h.index("Kernel Name")]
        f.write(f"\n== {name}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}\n")
        vals = {}
        for i, c in enumerate(h):
            if WANT.search(c):
                f.write(f"{c:95s} {r[i]:>16s} {units[i]}\n")
                vals[c] = r[i]
        try:
            rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
            ur, uw = units[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_write.sum")]
            f.write(f"dram traffic per launch: read {rd} {ur} + write {wr} {uw}\n")
        except Exception:
            pass

src = run(["--page", "source", "--csv"])
blocks = re.split(r'(?m)^"Kernel Name",', src)
for blk in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
    kname = rows[0][1]
    short = re.sub(r"[^A-Za-z0-9_]+", "_", kname.split("(")[0].split("<")[0].replace("void ", "")).strip("_")
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hh = rows[hi]
    iS, iE, iN = hh.index("Source"), hh.index("Instructions Executed"), hh.index("# Samples")
    stall_cols = [i for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
    ops, samples, stalls, seen, total = Counter(), Counter(), Counter(), set(), 0
    for r in rows[hi + 1:]:
        if len(r) <= iE or not r[0].startswith("0x") or r[0] in seen:
            continue
        seen.add(r[0])
        toks = r[iS].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).rstrip(";")
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDG", "STG"):
            base = ".".join([base] + [p for p in op.split(".")[1:] if p in ("64", "128")])
        e, n = int(float(r[iE] or 0)), int(float(r[iN] or 0))
        ops[base] += e; samples[base] += n; total += e
        for i in stall_cols:
            stalls[hh[i][6:]] += int(float(r[i] or 0))
    with open(f"{prefix}_{short}_mix.txt", "w") as f:
        f.write(f"# {kname}\n# executed warp-instructions by opcode (source page), total {total}\n")
        for k, v in ops.most_common(24):
            f.write(f"{k:12s} {v:12d} {100.0 * v / max(total, 1):6.2f}%   stall samples {samples[k]}\n")
        ts = sum(stalls.values()) or 1
        f.write("stall samples: " + ", ".join(f"{k} {100.0 * v / ts:.1f}%" for k, v in stalls.most_common(10)) + "\n")
print("wrote", prefix + "_metrics.txt and per-kernel mixes")
