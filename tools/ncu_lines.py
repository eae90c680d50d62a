#!/usr/bin/env python3
"""Top source lines of an ncu report by warp instructions executed (and their share of samples)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r): d.setdefault(k, v)
        def num(k):
            try: return int(d.get(k, "0") or 0)
            except ValueError: return 0
        data.append((num("Instructions Executed"), num("# Samples"), cur, int(r[0]), r[1].strip(), num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive")))
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print(f"total warp instructions {ti}")
for n, s, f, ln, src, w, we in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{100.0 * n / ti:5.1f}% inst {100.0 * s / ts:5.1f}% smp  wf {w:9d} (+{we:9d})  {f}:{ln:<5d} {src[:100]}")
