"""Where one batch-of-one fit spends its time (BASELINE.json configs[1] / configs[3]): wall clock of avb_fit, device time per
kernel class, and the CTA-time split of the flow kernel (at batch 1 the solve runs on ONE CTA, so its CTA ms is critical path).
    python tools/single_fit_profile.py [--jtj fp64|tensor] [--frames 4]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jtj", default="fp64")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--beta-pose", type=float, default=None, help="override beta_pose (0 switches the pose prior off)")
    args = ap.parse_args()
    from avatar_b200 import Fitter, default_options, _lib
    model, pr = bench.make_model()
    part_map, num_parts = pr["part_map"], int(pr["num_parts"])
    xg, x0 = bench.gen_params(model, range(args.frames))
    pf = Fitter(model, num_parts, part_map, args.frames, 16, 0)
    clouds_gt, _, _ = pf.avatar_update(xg)
    pf.close()
    pts, labs, off = bench.render_frames(model, part_map, clouds_gt)
    ft = Fitter(model, num_parts, part_map, 1, int(max(len(p) for p in pts)) + 64, 0)
    opt = default_options()
    opt.function_tolerance = 0.0
    if args.beta_pose is not None:
        opt.beta_pose = args.beta_pose
    opt.jtj_precision = {"fp64": _lib.JTJ_FP64, "tensor": _lib.JTJ_BF16_TENSOR}[args.jtj]
    out = {"jtj": args.jtj, "points": [len(p) for p in pts]}
    wall = []
    for r in range(5):
        for i in range(args.frames):
            tA = time.perf_counter()
            ft.fit_batch(pts[i], labs[i], np.array([0, len(pts[i])]), x0[i][None], opt)
            wall.append(time.perf_counter() - tA)
    out["fit_wall_ms_median"] = 1e3 * float(np.median(wall[args.frames:]))
    # resident: upload once, time fit_resident + sync
    ft.upload(pts[0], labs[0], np.array([0, len(pts[0])]))
    res = []
    for r in range(8):
        ft.timer_start()
        ft.fit_resident(x0[0][None], opt)
        res.append(ft.timer_stop())
    out["fit_resident_device_ms_median"] = float(np.median(res[2:]))
    ft.set_profiling(True)
    ft.fit_resident(x0[0][None], opt)
    ft.synchronize()
    out["kernel_ms"] = {k: round(v[0], 4) for k, v in ft.kernel_ms().items()}
    out["flow_task_cta_ms"] = {k: round(v, 4) for k, v in ft.flow_task_ms().items()}
    out["flow_phase_cta_ms"] = {k: round(v, 4) for k, v in ft.flow_phase_ms().items()}
    ft.set_profiling(False)
    ft.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
