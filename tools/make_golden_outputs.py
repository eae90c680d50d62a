#!/usr/bin/env python3
"""Freeze golden OUTPUT vectors of the CPU oracle into tests/golden/expected_r1.npz (the reference has no
golden vectors of its own, so the build freezes and versions its own).  Inputs are the committed fixtures
(model_synth.npz, prior_synth.npz) and seeded synthetic frames; everything here runs on the CPU.

    python tools/make_golden_outputs.py          # rewrites tests/golden/expected_r1.npz
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def golden_cases():
    """the seeded inputs every golden entry is computed from (shared with the tests)"""
    import oracle as orc
    from avatar_b200 import AvatarModel, GaussianMixture
    from harness import synth
    pr = np.load(os.path.join(GOLD, "prior_synth.npz"))
    g = GaussianMixture.from_arrays(pr["weights"], pr["means"], pr["covs"])
    model = AvatarModel(npz_path=os.path.join(GOLD, "model_synth.npz"), pose_prior=g)
    om = orc.OracleModel(os.path.join(GOLD, "model_synth.npz"), pr)
    oo = orc.OracleOptimizer(om, int(pr["num_parts"]), pr["part_map"])
    frames = []
    for seed in range(3):
        rng = np.random.default_rng(1000 + seed)
        x_gt = synth.random_params(model, rng)
        x0 = synth.perturbed_start(model, x_gt, rng)
        cloud_gt, _, _ = om.update_x(x_gt)
        pts, lab, depth, part = synth.render_cloud(model, cloud_gt, pr["part_map"])
        frames.append(dict(x_gt=x_gt, x0=x0, pts=pts, lab=lab, depth=depth, part=part, cloud_gt=cloud_gt))
    return orc, model, pr, om, oo, frames


def compute():
    from harness import synth
    orc, model, pr, om, oo, frames = golden_cases()
    nparts = int(pr["num_parts"])
    out = {}
    opt = orc.default_options(orc.SOLVER_GN_LM)
    opt.function_tolerance = 0.0
    tree = synth.random_rtree(np.random.default_rng(21), nparts)
    vp = synth.vertex_parts(model, pr["part_map"])
    faces = np.ascontiguousarray(model.mesh, dtype=np.int32)
    intrin = (synth.FX, synth.CX, synth.FY, synth.CY)
    for i, fr in enumerate(frames):
        x, st, _, nn = oo.optimize(fr["pts"], fr["lab"], fr["x0"], opt)
        out[f"fit_x_{i}"] = x
        out[f"fit_cost_{i}"] = np.array([st.initial_cost, st.final_cost])
        out[f"fit_iters_{i}"] = np.array([st.iterations, st.accepted_steps, st.num_correspondences], np.int64)
        out[f"nn_crc_{i}"] = crc(nn.astype(np.int32))
        out[f"num_points_{i}"] = np.int64(len(fr["pts"]))
        out[f"cloud_crc_{i}"] = crc(fr["pts"])                     # the cloud the harness back-projected (== orc_build_cloud)
        p2, l2 = orc.build_cloud(fr["depth"], fr["part"], intrin, nparts, None, 2)
        out[f"cloud_stride2_crc_{i}"] = np.array([crc(p2), crc(l2)], np.uint32)
        lab_img = orc.rtree_predict(fr["depth"], tree, None, 2, True)
        out[f"rtree_crc_{i}"] = crc(lab_img)
        r = orc.render(fr["cloud_gt"], faces, vp, synth.WIDTH, synth.HEIGHT, intrin)
        out[f"render_crc_{i}"] = np.array([crc(r["depth"]), crc(r["parts"]), crc(r["faces"])], np.uint32)
        out[f"render_px_{i}"] = np.array([(r["depth"] > 0).sum(), (r["parts"] != 255).sum()], np.int64)
    return out


if __name__ == "__main__":
    res = compute()
    np.savez_compressed(os.path.join(GOLD, "expected_r1.npz"), **res)
    print("wrote", os.path.join(GOLD, "expected_r1.npz"), len(res), "entries")
