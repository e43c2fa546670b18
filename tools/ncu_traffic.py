#!/usr/bin/env python
"""DRAM traffic per kernel launch from an ncu --set full report -> JSON (bench.py reads it for roofline.traffic).
  python tools/ncu_traffic.py gpurun_out/x.ncu-rep "<command the report was captured from>" > profiles/rNN_traffic.json"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1].split("<")[0].strip()
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[k]]) * scale[units[col[k]]]
    res.setdefault(short, []).append(tot)
print(json.dumps({"source": sys.argv[1], "command": sys.argv[2] if len(sys.argv) > 2 else None,
                  "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none",
                  "traffic_bytes_per_launch": {k: sum(v) / len(v) for k, v in res.items()}}, indent=1))
