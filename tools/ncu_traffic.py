#!/usr/bin/env python
"""Summarise an `ncu --set full` capture: profiles/<tag>_<kernel>_full_raw.csv (the raw page) and
profiles/ncu_traffic.json (DRAM bytes per launch of each captured kernel, read by bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/r1i_k_solve_velocity.ncu-rep r1i 16384 pyramid
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag, worlds, workload = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
try:
    out = json.load(open(path))
except Exception:
    out = {}
out = {k: v for k, v in out.items() if isinstance(v, dict)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    out[name] = {"source": f"profiles/{tag}_{name}_full_raw.csv (ncu --set full --clock-control none)", "worlds": worlds,
                 "workload": workload, "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                 "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]) * (1e-3 if u["gpu__time_duration.sum"] == "us" else 1.0),
                 "registers": int(float(d["launch__registers_per_thread"])), "grid": int(float(d["launch__grid_size"]))}
    with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_full_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(hdr); w.writerow(units); w.writerow(r)
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
