#!/usr/bin/env python3
"""Top SASS instructions by stall samples from `ncu --page source --csv`.  usage: ncu_hot.py rep [kernel-index] [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
s = sections[which]
idx = {h: i for i, h in enumerate(s["hdr"])}
data = s["rows"]
tot = sum(int(r[idx["# Samples"]]) for r in data)
ti = sum(int(r[idx["Instructions Executed"]]) for r in data)
print(f"{len(sections)} kernels; [{which}] {s['name'][:80]}: samples {tot}, warp inst {ti}, sass lines {len(data)}")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    print(r[idx["# Samples"]].rjust(7), r[idx["Instructions Executed"]].rjust(11), r[idx["Avg. Threads Executed"]].rjust(6), " ", r[idx["Source"]][:110])
