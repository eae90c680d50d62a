"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/avatar_b200.h
declares, validates its inputs, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(build_all):
    from avatar_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "avatar_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(avb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(_lib.lib, name), name


def test_default_options_match_reference_defaults(build_all):
    from avatar_b200 import default_options
    o = default_options()
    # include/AvatarOptimizer.h:19,27,33,36,39 and AvatarOptimizer.cpp:1333
    assert (o.icp_iters, o.max_iters_per_icp, o.nn_step, o.enable_occlusion) == (1, 10, 20, 1)
    assert (o.beta_pose, o.beta_shape, o.function_tolerance) == (0.1, 1.0, 1e-4)


def test_model_create_validates(model):
    from avatar_b200 import _lib
    assert model.handle is not None
    V, J, K, F = (C.c_int32() for _ in range(4))
    _lib.check(_lib.lib.avb_model_dims(model.handle, C.byref(V), C.byref(J), C.byref(K), C.byref(F)))
    assert (V.value, J.value, K.value, F.value) == (6890, 24, 10, 13776)
    assert _lib.lib.avb_param_dim(model.handle) == 109 and _lib.lib.avb_tangent_dim(model.handle) == 85
    d = _lib.ModelDesc()
    h = C.c_void_p()
    assert _lib.lib.avb_model_create(C.byref(d), C.byref(h)) == 1  # AVB_ERR_INVALID
    assert b"dimension" in _lib.lib.avb_last_error()


def test_model_prior_matches_oracle(model, omodel):
    """GaussianMixture::load maths (GaussianMixture.cpp:22-76) in the library vs the oracle restatement"""
    from avatar_b200 import _lib
    pc, cl = np.zeros((8, 69, 69)), np.zeros(8)
    _lib.check(_lib.lib.avb_model_get_prior(model.handle, _lib.ptr(pc), _lib.ptr(cl)))
    opc, ocl = omodel.prior()
    np.testing.assert_allclose(cl, ocl, rtol=1e-13)
    np.testing.assert_allclose(pc, opc, rtol=0, atol=1e-9)


def test_no_cpu_fallback(model, prior_arrays):
    """without a usable GPU every compute entry point must fail loudly"""
    from avatar_b200 import _lib, Fitter, AvbError
    if _lib.lib.avb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(AvbError) as e:
        Fitter(model, int(prior_arrays["num_parts"]), prior_arrays["part_map"], 1, 1024)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_synth_render_and_backproject(model, omodel, prior_arrays):
    from harness import synth
    rng = np.random.default_rng(0)
    x = synth.random_params(model, rng)
    cloud, _, _ = omodel.update_x(x)
    pts, lab, depth, part = synth.render_cloud(model, cloud, prior_arrays["part_map"])
    assert depth.shape == (576, 640) and 4000 < len(pts) < 60000
    assert lab.min() >= 0 and lab.max() < 16
    # every back-projected point re-projects to its pixel: (c - cx) z / fx, -(r - cy) z / fy, z
    r, c = np.nonzero(depth > 0)
    np.testing.assert_allclose(pts[:, 0] * synth.FX / pts[:, 2] + synth.CX, c, atol=1e-3)
    np.testing.assert_allclose(-pts[:, 1] * synth.FY / pts[:, 2] + synth.CY, r, atol=1e-3)
    # points lie on the posed surface: nearest model vertex within a few cm
    d = np.sqrt(((pts[::97, None, :] - cloud[None, :, :]) ** 2).sum(-1)).min(axis=1)
    assert d.max() < 0.06
    sub, _ = synth.backproject(depth, part, interval=12)
    assert 0 < len(sub) < len(pts) / 100


def test_shard_and_gather_world2_gloo(tmp_path):
    """frames shard across ranks with no data-path collective; one gather restores global order"""
    import subprocess
    import sys
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), script, str(tmp_path / "out.npy")]
    subprocess.run(cmd, check=True, timeout=300, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    out = np.load(str(tmp_path / "out.npy"))
    assert out.shape == (10, 109)
    np.testing.assert_array_equal(out[:, 0], np.arange(10))


def test_widened_entry_points_validate_without_a_gpu(build_all):
    """the entry points of the widened rows (cloud construction, RTree, renderer) reject null handles / arguments with
    AVB_ERR_INVALID before touching CUDA"""
    from avatar_b200 import _lib
    lib = _lib.lib
    img = _lib.ImageDesc(640, 576, 504.0, 320.0, 504.0, 288.0, 1, 16, 2)
    view = _lib.RenderDesc(640, 576, 504.0, 320.0, 504.0, 288.0)
    tree = _lib.RTreeDesc()
    ms = (C.c_float * 4)()
    calls = [
        lambda: lib.avb_upload_depth_batch(None, 1, None, None, None, C.byref(img), None),
        lambda: lib.avb_download_batch(None, None, None, None),
        lambda: lib.avb_last_cloud_ms(None, ms),
        lambda: lib.avb_fitter_set_rtree(None, C.byref(tree)),
        lambda: lib.avb_rtree_predict_batch(None, 1, None, 640, 576, None, 1, 1, None),
        lambda: lib.avb_last_rtree_ms(None, ms),
        lambda: lib.avb_render_batch(None, 1, None, C.byref(view), None, None, None),
        lambda: lib.avb_last_render_ms(None, ms),
    ]
    for call in calls:
        assert call() == 1                                  # AVB_ERR_INVALID
        assert len(lib.avb_last_error()) > 0
