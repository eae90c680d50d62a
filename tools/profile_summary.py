#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep into the tracked summaries under profiles/:
   r1_ncu_<kernel>_raw.csv   (metric, unit, value) of one representative launch
   r1_ncu_<kernel>_hot_lines.txt   source lines by warp-stall samples (per kernel)
usage: tools/profile_summary.py [--round r2] [--label fp64] gpurun_out/r1_front.ncu-rep gpurun_out/r1_eval.ncu-rep ...
   --round: file-name prefix (default r1); --label: appended to the kernel tag (e.g. flow -> flow_fp64)"""
import csv, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"gpu__time_duration|dram__bytes|dram__throughput|gpu__dram_throughput|sm__throughput|sm__warps_active|"
                  r"pipe_fp64|ops_path_tensor_src_fp64|pipe_tensor|smsp__issue_active|launch__|bank_conflicts|lts__t_sector_hit_rate|"
                  r"l1tex__t_sector_hit_rate|smsp__warp_issue_stalled|smsp__average_warps_issue_stalled|sm__inst_executed_pipe|"
                  r"lts__throughput|l1tex__throughput|smsp__inst_executed.sum|sm__cycles_active.avg")
argv = sys.argv[1:]
RND, LABEL = "r1", ""
while argv and argv[0].startswith("--"):
    if argv[0] == "--round":
        RND = argv[1]
    elif argv[0] == "--label":
        LABEL = "_" + argv[1]
    argv = argv[2:]
seen = {}
for rep in argv:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kidx = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = re.match(r"(?:void\s+)?(?:\w+::)*(\w+)", r[kidx]).group(1)   # "void avb::lm_flow_kernel<(bool)0, (int)3>(...)" -> lm_flow_kernel
        tag = re.sub(r"<.*", "", name.split("::")[-1]).replace("lm_", "").replace("_kernel", "") + LABEL   # templates: drop <...>
        n = seen.get(tag, 0)
        seen[tag] = n + 1
        suffix = "" if n == 0 else f"_{n + 1}"
        path = os.path.join(root, "profiles", f"{RND}_ncu_{tag}{suffix}_kernel_raw.csv")
        with open(path, "w") as fh:
            fh.write(f"# {name}, one launch, ncu --set full --clock-control none ({os.path.basename(rep)})\n")
            for h, u, v in zip(hdr, units, r):
                if KEEP.search(h):
                    fh.write(f"{h},{u},{v}\n")
        print("wrote", path)
    # hot lines per kernel of this report
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur_kernel, cur_file, hdrl, data = None, None, None, {}
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "Function Name":
            cur_kernel = re.sub(r"<.*", "", r[1].split("(")[0].split("::")[-1])
        elif len(r) == 2 and r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdrl = r
        elif hdrl and len(r) == len(hdrl) and r[0].isdigit():
            d = {}
            for k, v in zip(hdrl, r):
                d.setdefault(k, v)
            try:
                s = int(d.get("# Samples", "0") or 0)
            except ValueError:
                s = 0
            data.setdefault(cur_kernel, []).append((s, cur_file, int(r[0]), r[1].strip(), d))
    for k, rows_k in data.items():
        if k is None:
            continue
        tag = k.replace("lm_", "").replace("_kernel", "") + LABEL
        tot = sum(x[0] for x in rows_k) or 1
        path = os.path.join(root, "profiles", f"{RND}_ncu_{tag}_kernel_hot_lines.txt")
        with open(path, "w") as fh:
            fh.write(f"# {k}: source lines by warp-stall samples (total {tot}); top stall reasons per line\n")
            for s, f, ln, src, d in sorted(rows_k, key=lambda t: -t[0])[:40]:
                st = {kk[6:]: int(v) for kk, v in d.items() if kk.startswith("stall_") and "Not Issued" not in kk and v not in ("", "0", "-")}
                main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
                fh.write(f"{100.0 * s / tot:5.1f}%  {f}:{ln:<5d} {src[:100]:100s} {main}\n")
        print("wrote", path)
