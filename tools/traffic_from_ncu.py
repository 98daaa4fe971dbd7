"""profiles/traffic.json from an ncu report of the bench step: dram__bytes_read.sum + dram__bytes_write.sum per launch of
the two step kernels, stamped with the git head the report was taken at (bench.py prints it as roofline.traffic).

    python tools/traffic_from_ncu.py gpurun_out/r2_prof.ncu-rep [git-head]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
head = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout)))
h, units = raw[0], raw[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in raw[2:]:
    name = r[h.index("Kernel Name")]
    tot = 0.0
    for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(col)
        tot += float(r[i]) * scale[units[i]]
    per.setdefault(name, []).append(tot)
out = {"git_head": head, "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches)",
       "kernels": {k: sum(v) / len(v) for k, v in per.items()}}
for k, v in out["kernels"].items():
    if "mask_istft_kernel" in k and k.rstrip(")").split("(")[0].rstrip(">").endswith(("1, 1", "true, true", "(bool)1, (bool)1")):
        out["mask_istft_feature_bytes_per_launch"] = v
    elif "mask_istft_kernel" in k and "mask_istft_feature_bytes_per_launch" not in out and ", 1, " in k:
        out["mask_istft_feature_bytes_per_launch"] = v
    elif "mask_istft_kernel" in k:
        out["mask_istft_bytes_per_launch"] = v
    elif "stft_kernel" in k and "istft" not in k:
        out["stft_dual_bytes_per_launch" if k.rstrip(")").split("(")[0].rstrip(">").endswith(("1", "true")) else "stft_bytes_per_launch"] = v
out["note"] = ("the feature-fed synthesis really reads the 4TN-byte linear spectrum (197 MB at C2) instead of the 4n-byte waveform: its DRAM traffic "
               "exceeds the 492.3 MB of algorithmic bytes the roofline fraction is quoted on (SURVEY 8d); the dual-output STFT writes 2 x 4TN")
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
