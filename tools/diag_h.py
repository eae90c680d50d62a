#!/usr/bin/env python3
"""Diagnostic: where does the tensor-path J^T J differ most from the oracle's (one frame, start point)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc
from avatar_b200 import AvatarModel, GaussianMixture, Fitter, default_options, _lib
from harness import synth
gold = os.path.join(ROOT, "tests", "golden")
pr = np.load(os.path.join(gold, "prior_synth.npz"))
g = GaussianMixture.from_arrays(pr["weights"], pr["means"], pr["covs"])
model = AvatarModel(npz_path=os.path.join(gold, "model_synth.npz"), pose_prior=g)
om = orc.OracleModel(os.path.join(gold, "model_synth.npz"), pr)
oo = orc.OracleOptimizer(om, int(pr["num_parts"]), pr["part_map"])
rng = np.random.default_rng(1000)
x_gt = synth.random_params(model, rng); x0 = synth.perturbed_start(model, x_gt, rng)
cloud_gt, _, _ = om.update_x(x_gt)
pts, lab, _, _ = synth.render_cloud(model, cloud_gt, pr["part_map"])
ft = Fitter(model, int(pr["num_parts"]), pr["part_map"], 1, len(pts) + 16)
off = np.array([0, len(pts)])
ft.upload(pts, lab, off)
for name, prec in (("tensor", _lib.JTJ_BF16_TENSOR), ("fp64", _lib.JTJ_FP64)):
    o = default_options(); o.jtj_precision = prec
    ft.debug_correspond(x0[None], o)
    nn = ft.debug_read(_lib.TAP_NN)
    cost, grad, H = ft.debug_evaluate(x0[None], o)
    oc, og, oH = oo.evaluate(x0, pts, nn, o.beta_pose, o.beta_shape)
    scale = np.sqrt(np.outer(np.diag(oH), np.diag(oH)))
    err = np.abs(H[0] - oH) / scale
    print(name, "cost rel", abs(cost[0] - oc) / oc, "grad rel", np.abs(grad[0] - og).max() / np.abs(og).max(), "H max", err.max())
    idx = np.dstack(np.unravel_index(np.argsort(-err, axis=None)[:12], err.shape))[0]
    for i, j in idx:
        print(f"   H[{i},{j}] = {H[0][i, j]:.9g}  oracle {oH[i, j]:.9g}  err/scale {err[i, j]:.2e}  diag {oH[i, i]:.4g} {oH[j, j]:.4g}")
    print("   diag rel errs:", np.array2string(np.abs(np.diag(H[0]) - np.diag(oH)) / np.diag(oH), precision=1, max_line_width=200))
