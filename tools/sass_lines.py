#!/usr/bin/env python3
"""Attribute the SASS instructions of one kernel to source lines (code-size hunting).
usage: tools/sass_lines.py <kernel-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "avatar_b200", "libavatar_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cnt = collections.Counter()
for f in os.listdir(tmp):
    if not f.endswith(".cubin") or "-" in f: continue
    out = subprocess.run(["nvdisasm", "--print-line-info", f], cwd=tmp, capture_output=True, text=True).stdout
    fn = line = None
    for l in out.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m: fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l) and fn and kern in fn: cnt[line] += 1
print("total", sum(cnt.values()))
srcs = {}
for (f, ln), c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    path = os.path.join(root, "avatar_b200", "csrc", f)
    if f not in srcs: srcs[f] = open(path).read().split("\n") if os.path.exists(path) else None
    txt = srcs[f][ln - 1].strip()[:100] if srcs[f] else ""
    print(c, f, ln, txt)
if len(sys.argv) > 3:
    # region totals: pairs "name:first-last" on avb_lm.cu
    for spec in sys.argv[3:]:
        name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
        print(name, sum(c for (f, ln), c in cnt.items() if f == "avb_lm.cu" and a <= ln <= b))
    for f in set(f for (f, _) in cnt if f != "avb_lm.cu"): print(f, sum(c for (g, _), c in cnt.items() if g == f))
