#!/usr/bin/env python3
"""Small driver for ncu captures of the image-side kernels (cloud construction, RTree prediction, renderer):
   ncu --set full --clock-control none --import-source on -k regex:'rtree_|render_|cloud_' -s 8 -c 8 -o gpurun_out/r1_images python tools/ncu_images.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatar_b200 import AvatarModel, GaussianMixture, Fitter  # noqa: E402
from harness import synth

GOLD = os.path.join(ROOT, "tests", "golden")
pr = np.load(os.path.join(GOLD, "prior_synth.npz"))
g = GaussianMixture.from_arrays(pr["weights"], pr["means"], pr["covs"])
model = AvatarModel(npz_path=os.path.join(GOLD, "model_synth.npz"), pose_prior=g)
nparts, part_map = int(pr["num_parts"]), pr["part_map"]
B = 64
xs = np.stack([synth.random_params(model, np.random.default_rng(100000 + s)) for s in range(B)])
ft = Fitter(model, nparts, part_map, B, B * 40000)
intrin = (synth.FX, synth.CX, synth.FY, synth.CY)
ft.set_rtree(synth.random_rtree(np.random.default_rng(7), nparts), nparts)
for _ in range(2):   # first pass = warm-up (skipped by -s), second pass is captured
    img = ft.render(xs, synth.WIDTH, synth.HEIGHT, intrin, want=("depth", "parts"))      # prepare, cover, resolve
    lab = ft.rtree_predict(img["depth"], None, 2, True)                                  # predict, upscale
    ft.upload_depth(img["depth"], img["parts"], intrin, nparts)                          # count, compact
    ft.synchronize()
print("rendered", int((img["depth"] > 0).sum()), "px; labels", int((lab != 255).sum()))
ft.close()
