#!/usr/bin/env python3
"""Aggregate an ncu source-page export by code region of avb_lm.cu (function bodies): warp instructions executed,
stall samples and shared-memory wavefronts per region.  usage: ncu_regions.py report.ncu-rep [file.cu marker=label ...]"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = open(os.path.join(ROOT, "avatar_b200", "csrc", "avb_lm.cu")).read().split("\n")
MARKS = [("queue", "__device__ void flow_push"), ("lm_prep", "lm_prep_kernel(DevModel"), ("rows_body", "__device__ void rows_body"),
         ("gram_body+emit", "constexpr int kGramThreads"), ("tc helpers", "#ifndef AVB_TC_TERMS"), ("emit_partial_tc", "__device__ void emit_partial_tc"),
         ("fused_body", "__device__ void fused_body"), ("cholesky", "constexpr int kNBsq"), ("back_solve", "__device__ void warp_back_solve"),
         ("solve_body", "__device__ bool solve_body"), ("flow loop", "constexpr int kFlowMinCtasTc"), ("launchers", "size_t lm_prep_smem")]
marks = []
for name, pat in MARKS:
    for i, l in enumerate(src):
        if pat in l:
            marks.append((i + 1, name))
            break
marks.sort()


def region(f, ln):
    if f != "avb_lm.cu":
        return f
    r = "head"
    for l, name in marks:
        if l <= ln:
            r = name
    return r


cur, hdr = None, None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)

        def num(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        a = agg.setdefault(region(cur, int(r[0])), [0, 0, 0, 0, 0, 0])
        a[0] += num("Instructions Executed")
        a[1] += num("# Samples")
        a[2] += num("L1 Wavefronts Shared")
        a[3] += num("L1 Wavefronts Shared Excessive")
        a[4] += num("stall_long_sb")
        a[5] += num("stall_barrier")
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"{'region':24s} {'inst%':>6s} {'samples%':>8s} {'smem wavefronts':>16s} {'excess':>10s} {'long_sb%':>8s} {'barrier%':>8s}   (warp instructions {ti})")
for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{k:24s} {100 * a[0] / ti:6.1f} {100 * a[1] / ts:8.1f} {a[2]:16d} {a[3]:10d} {100 * a[4] / ts:8.1f} {100 * a[5] / ts:8.1f}")
