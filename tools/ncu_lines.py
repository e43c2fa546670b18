#!/usr/bin/env python
"""Aggregate an .ncu-rep source page (needs -lineinfo + --import-source on) by CUDA source line.
  python tools/ncu_lines.py x.ncu-rep [kernel-regex] [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, cur, curfile = collections.OrderedDict(), None, None
for r in rows:
    if r and r[0] == "File Name":
        curfile = r[1].split("/")[-1]; continue
    if len(r) <= max(ie, ss): continue
    if r[0] not in ("", "Line No"):
        cur = (curfile, r[0], r[1].strip()[:110]); agg.setdefault(cur, [0, 0]); continue
    if r[0] == "" and cur and r[2] not in ("...", "-"):
        try:
            agg[cur][0] += int(r[ie]); agg[cur][1] += int(r[ss])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print("warp instructions", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / tots))[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (v[0] * 100 / tot, v[1] * 100 / tots, k[0], k[1], k[2]))
