#!/usr/bin/env python3
"""Summarise an ncu --page source --csv --print-source cuda export: top source lines by warp-stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
hdr = None
data = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            s = int(d.get("# Samples", "0") or 0)
        except ValueError:
            s = 0
        data.append((s, cur, int(r[0]), r[1].strip(), d))
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}")
for s, f, ln, src, d in sorted(data, key=lambda t: -t[0])[:top]:
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0", "-")}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100.0 * s / tot:5.1f}%  {f}:{ln:<5d} {src[:90]:90s} {main}")
