"""ctypes wrapper of oracle/libavatar_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (avatar_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavatar_oracle.so")
REF_NANOFLANN_PATH = os.path.join(_HERE, "_ref", "libref_nanoflann.so")

SOLVER_BFGS_WOLFE, SOLVER_GN_LM = 0, 1


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


class Options(C.Structure):
    _fields_ = [("icp_iters", C.c_int32), ("max_iters_per_icp", C.c_int32), ("beta_pose", C.c_double),
                ("beta_shape", C.c_double), ("enable_occlusion", C.c_int32), ("solver", C.c_int32),
                ("function_tolerance", C.c_double), ("num_threads", C.c_int32), ("nn_method", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("num_correspondences", C.c_int32), ("iterations", C.c_int32), ("evaluations", C.c_int32),
                ("accepted_steps", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("seconds", C.c_double)]


def default_options(solver=SOLVER_GN_LM):
    return Options(1, 10, 0.1, 1.0, 1, solver, 1e-4, 4, 1)


if not os.path.exists(LIB_PATH):
    build()
_lib = C.CDLL(LIB_PATH)
_P = C.c_void_p
_sig = {
    "orc_model_create": (_P, [C.c_int] * 4 + [_P] * 6),
    "orc_model_destroy": (None, [_P]),
    "orc_model_set_prior": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "orc_model_load_prior_text": (C.c_int, [_P, C.c_char_p]),
    "orc_model_get_joint_reg": (None, [_P, _P, _P, _P]),
    "orc_model_get_assigned": (C.c_int, [_P, _P, _P, _P]),
    "orc_model_get_prior": (None, [_P, _P, _P]),
    "orc_avatar_update": (None, [_P] * 7),
    "orc_rotmat_to_quat": (None, [_P, _P]),
    "orc_quat_to_rotmat": (None, [_P, _P]),
    "orc_gmm_residual": (C.c_int, [_P, _P, _P]),
    "orc_optimizer_create": (_P, [_P, C.c_int, _P]),
    "orc_optimizer_destroy": (None, [_P]),
    "orc_visibility": (None, [_P, _P, _P]),
    "orc_find_nn": (None, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "orc_evaluate": (C.c_double, [_P, _P, _P, _P, C.c_int, C.c_double, C.c_double, C.c_int, _P, _P]),
    "orc_vertex_jacobian": (None, [_P, _P, C.c_int, _P, _P]),
    "orc_vertex_position_chain": (None, [_P, _P, C.c_int, _P]),
    "orc_retract": (None, [_P, _P, _P, _P]),
    "orc_optimize": (C.c_int, [_P, _P, _P, C.c_int, _P, C.POINTER(Options), C.POINTER(Stats), _P, C.c_int,
                               C.POINTER(C.c_int), _P]),
    "orc_build_cloud": (C.c_int64, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, C.c_int64]),
    "orc_rtree_predict": (None, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "orc_render": (None, [_P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "orc_render_prelude": (None, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _P]),
    "orc_render_lambert": (None, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "orc_param_dim": (C.c_int, [_P]),
    "orc_tangent_dim": (C.c_int, [_P]),
}
for _n, (_r, _a) in _sig.items():
    getattr(_lib, _n).restype = _r
    getattr(_lib, _n).argtypes = _a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleModel:
    """ark::AvatarModel restated (npz branch)."""

    def __init__(self, npz_path, prior=None):
        z = np.load(npz_path)
        vt, sd = _f64(z["v_template"]), _f64(z["shapedirs"])
        jr, wt = _f64(z["J_regressor"]), _f64(z["weights"])
        parents = np.ascontiguousarray(z["kintree_table"][0].astype(np.uint32).astype(np.int32))
        faces = np.ascontiguousarray(np.asarray(z["f"]).astype(np.int64).astype(np.int32))
        self.V, self.J, self.K, self.F = vt.shape[0], parents.shape[0], sd.shape[2], faces.shape[0]
        self.faces, self.parents = faces, parents
        self.h = _lib.orc_model_create(self.V, self.J, self.K, self.F, _p(vt), _p(sd), _p(jr), _p(wt), _p(parents),
                                       _p(faces))
        self.C = self.D = 0
        if prior is not None:
            w, mu, cov = (_f64(prior[k]) for k in ("weights", "means", "covs"))
            self.C, self.D = mu.shape
            assert _lib.orc_model_set_prior(self.h, self.C, self.D, _p(w), _p(mu), _p(cov)) == 0

    def load_prior_text(self, path, C_, D_):
        rc = _lib.orc_model_load_prior_text(self.h, path.encode())
        self.C, self.D = C_, D_
        return rc

    def joint_reg(self):
        base, reg, init = np.zeros(3 * self.J), np.zeros((3 * self.J, self.K)), np.zeros(3 * self.J)
        _lib.orc_model_get_joint_reg(self.h, _p(base), _p(reg), _p(init))
        return base, reg, init

    def assigned(self):
        n = _lib.orc_model_get_assigned(self.h, None, None, None)
        start, joint, weight = np.zeros(self.V + 1, np.int32), np.zeros(n, np.int32), np.zeros(n)
        _lib.orc_model_get_assigned(self.h, _p(start), _p(joint), _p(weight))
        return start, joint, weight

    def prior(self):
        pc, cl = np.zeros((self.C, self.D, self.D)), np.zeros(self.C)
        _lib.orc_model_get_prior(self.h, _p(pc), _p(cl))
        return pc, cl

    def gmm_residual(self, x):
        out = np.zeros(self.D + 1)
        comp = _lib.orc_gmm_residual(self.h, _p(_f64(x)), _p(out))
        return out, comp

    def update(self, p, R, w):
        """Avatar::update: R (J,3,3). returns cloud (V,3), jointPos (J,3), jointTrans (J,12)"""
        cloud, jp, jt = np.zeros((self.V, 3)), np.zeros((self.J, 3)), np.zeros((self.J, 12))
        _lib.orc_avatar_update(self.h, _p(_f64(p)), _p(_f64(R)), _p(_f64(w)), _p(cloud), _p(jp), _p(jt))
        return cloud, jp, jt

    def update_x(self, x):
        J = self.J
        R = np.stack([quat_to_rotmat(x[3 + 4 * j:7 + 4 * j]) for j in range(J)])
        return self.update(x[:3], R, x[3 + 4 * J:])


def rotmat_to_quat(R):
    q = np.zeros(4)
    _lib.orc_rotmat_to_quat(_p(_f64(R)), _p(q))
    return q


def quat_to_rotmat(q):
    R = np.zeros((3, 3))
    _lib.orc_quat_to_rotmat(_p(_f64(q)), _p(R))
    return R


class OracleOptimizer:
    """ark::AvatarOptimizer restated."""

    def __init__(self, model, num_parts, part_map):
        self.model = model
        self.part_map = np.ascontiguousarray(part_map, dtype=np.int32)
        self.h = _lib.orc_optimizer_create(model.h, num_parts, _p(self.part_map))
        self.nx, self.P = _lib.orc_param_dim(self.h), _lib.orc_tangent_dim(self.h)

    def visibility(self, cloud):
        vis = np.zeros(self.model.V, dtype=np.uint8)
        _lib.orc_visibility(self.h, _p(_f64(cloud)), _p(vis))
        return vis

    def find_nn(self, cloud, vis, data, labels, method=0):
        data, labels = _f64(data), np.ascontiguousarray(labels, dtype=np.int32)
        idx = np.zeros(data.shape[0], dtype=np.int32)
        _lib.orc_find_nn(self.h, _p(_f64(cloud)), _p(np.ascontiguousarray(vis, dtype=np.uint8)), _p(data), _p(labels),
                         data.shape[0], method, _p(idx))
        return idx

    def evaluate(self, x, data, idx, beta_pose=0.1, beta_shape=1.0, want_H=True, num_threads=1):
        data, idx = _f64(data), np.ascontiguousarray(idx, dtype=np.int32)
        grad = np.zeros(self.P)
        H = np.zeros((self.P, self.P)) if want_H else None
        cost = _lib.orc_evaluate(self.h, _p(_f64(x)), _p(data), _p(idx), data.shape[0], beta_pose, beta_shape,
                                 num_threads, _p(grad), _p(H))
        return cost, grad, H

    def vertex_jacobian(self, x, v):
        pos, jac = np.zeros(3), np.zeros((3, self.P))
        _lib.orc_vertex_jacobian(self.h, _p(_f64(x)), v, _p(pos), _p(jac))
        return pos, jac

    def vertex_position_chain(self, x, v):
        pos = np.zeros(3)
        _lib.orc_vertex_position_chain(self.h, _p(_f64(x)), v, _p(pos))
        return pos

    def retract(self, x, delta):
        xp = np.zeros(self.nx)
        _lib.orc_retract(self.h, _p(_f64(x)), _p(_f64(delta)), _p(xp))
        return xp

    def optimize(self, data, labels, x, opt, trace_cap=0):
        data, labels = _f64(data), np.ascontiguousarray(labels, dtype=np.int32)
        x = np.array(x, dtype=np.float64, copy=True)
        st = Stats()
        trace = np.zeros((max(trace_cap, 1), self.nx))
        tl = C.c_int(0)
        nn = np.zeros(data.shape[0], dtype=np.int32)
        _lib.orc_optimize(self.h, _p(data), _p(labels), data.shape[0], _p(x), C.byref(opt), C.byref(st),
                          _p(trace) if trace_cap else None, trace_cap, C.byref(tl), _p(nn))
        return x, st, trace[:tl.value], nn


def build_cloud(depth, parts, intrin, num_parts, roi=None, interval=1):
    """demo.cpp:215-250 + Calibration.cpp:83-95 restated (orc_build_cloud): (points [N,3], labels [N]); raises
    ValueError on a label >= num_parts (where the reference exits)"""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    parts = np.ascontiguousarray(parts, dtype=np.uint8)
    h, w = depth.shape
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    r = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32)
    cap = int((parts != 255).sum()) + 1
    pts = np.zeros((cap, 3))
    lab = np.zeros(cap, dtype=np.int32)
    n = _lib.orc_build_cloud(_p(depth), _p(parts), w, h, _p(k), _p(r), int(interval), int(num_parts), _p(pts), _p(lab), cap)
    if n == -1:
        raise ValueError("body-part label >= num_parts")
    assert n >= 0
    return pts[:n].copy(), lab[:n].copy()


def rtree_predict(depth, tree, roi=None, interval=1, fill_in_gaps=True):
    """RTree::predictBest on one image (RTree.cpp:3184-3262) + upscaleGrid; tree: dict of u, v, thresh, lnode, rnode,
    leafid, leaf_best arrays (avatar_b200.synth.random_rtree layout)"""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    h, w = depth.shape
    t = {k: np.ascontiguousarray(tree[k], dtype=d) for k, d in (("u", np.float32), ("v", np.float32), ("thresh", np.float32),
                                                              ("lnode", np.int32), ("rnode", np.int32), ("leafid", np.int32),
                                                              ("leaf_best", np.uint8))}
    r = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32)
    out = np.zeros((h, w), dtype=np.uint8)
    _lib.orc_rtree_predict(_p(depth), w, h, len(t["thresh"]), _p(t["u"]), _p(t["v"]), _p(t["thresh"]), _p(t["lnode"]),
                           _p(t["rnode"]), _p(t["leafid"]), _p(t["leaf_best"]), _p(r), int(interval), int(bool(fill_in_gaps)), _p(out))
    return out


def render(cloud, faces, vertex_part, width, height, intrin, want=("depth", "parts", "faces")):
    """AvatarRenderer::renderDepth / renderPartMask / renderFaces restated (orc_render): dict of images + 'order'"""
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vp = np.ascontiguousarray(vertex_part, dtype=np.int32)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    depth = np.zeros((height, width), np.float32) if "depth" in want else None
    parts = np.zeros((height, width), np.uint8) if "parts" in want else None
    fids = np.zeros((height, width), np.int32) if "faces" in want else None
    order = np.zeros(faces.shape[0], np.int32)
    _lib.orc_render(_p(cloud), cloud.shape[0], _p(faces), faces.shape[0], _p(vp), width, height, _p(k), _p(depth), _p(parts),
                    _p(fids), _p(order))
    return dict(depth=depth, parts=parts, faces=fids, order=order)


def render_lambert(cloud, faces, width, height, intrin, taps=False):
    """AvatarRenderer::renderLambert restated (orc_render_lambert): uint8 image [H, W]; taps=True also returns the per-vertex
    lambert values [V] and the visibility of the ordered faces [F]"""
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    gray = np.zeros((height, width), np.uint8)
    lam = np.zeros(cloud.shape[0], np.float32)
    vis = np.zeros(faces.shape[0], np.uint8)
    _lib.orc_render_lambert(_p(cloud), cloud.shape[0], _p(faces), faces.shape[0], width, height, _p(k), _p(gray), _p(lam), _p(vis))
    return (gray, lam, vis) if taps else gray


def ref_paint_lambert(proj, faces_ordered, visible, lambert, width, height):
    """renderLambert's painting loop through the REFERENCE'S OWN paintTriangleBary<uint8_t> (oracle/_ref); None if not built"""
    if not os.path.exists(REF_PAINTERS_PATH):
        return None
    lib = C.CDLL(REF_PAINTERS_PATH)
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    fo = np.ascontiguousarray(faces_ordered, dtype=np.int32)
    vis = np.ascontiguousarray(visible, dtype=np.uint8)
    lam = np.ascontiguousarray(lambert, dtype=np.float32)
    gray = np.zeros((height, width), np.uint8)
    lib.ref_paint_lambert.argtypes = [_P, C.c_int, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P]
    lib.ref_paint_lambert.restype = None
    lib.ref_paint_lambert(_p(proj), proj.shape[0], _p(fo), fo.shape[0], _p(vis), _p(lam), width, height, _p(gray))
    return gray


REF_PAINTERS_PATH = os.path.join(_HERE, "_ref", "libref_painters.so")


def render_prelude(cloud, faces, intrin):
    """getProjectedPoints + getOrderedFaces + the grazing test restated: (proj [V,2] float32, order [F], grazing [F])"""
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    V, F = cloud.shape[0], faces.shape[0]
    proj = np.zeros((V, 2), np.float32)
    order = np.zeros(F, np.int32)
    grazing = np.zeros(F, np.uint8)
    _lib.orc_render_prelude(_p(cloud), V, _p(faces), F, _p(k), _p(proj), _p(order), _p(grazing))
    return proj, order, grazing


def ref_painters_render(cloud, faces, vertex_part, width, height, intrin):
    """the images painted by the REFERENCE'S OWN painters (AvatarHelpers.cpp compiled into oracle/_ref against the
    container stand-ins of oracle/shim), driven face by face in the renderer's order; None if oracle/_ref is not built"""
    if not os.path.exists(REF_PAINTERS_PATH):
        return None
    lib = C.CDLL(REF_PAINTERS_PATH)
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vp = np.ascontiguousarray(vertex_part, dtype=np.int32)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    V, F = cloud.shape[0], faces.shape[0]
    proj = np.zeros((V, 2), np.float32)
    order = np.zeros(F, np.int32)
    grazing = np.zeros(F, np.uint8)
    _lib.orc_render_prelude(_p(cloud), V, _p(faces), F, _p(k), _p(proj), _p(order), _p(grazing))
    of = np.ascontiguousarray(faces[order])
    zv = np.ascontiguousarray(cloud[of, 2].astype(np.float32))
    depth = np.zeros((height, width), np.float32)
    parts = np.full((height, width), 255, np.uint8)
    fids = np.full((height, width), -1, np.int32)
    lib.ref_paint_faces.argtypes = [_P, C.c_int, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P]
    lib.ref_paint_faces(_p(proj), V, _p(of), F, _p(grazing), _p(zv), _p(vp), width, height, _p(depth), _p(parts), _p(fids))
    return dict(depth=depth, parts=parts, faces=fids, order=order)


def ref_depth_to_xyz(depth, intrin):
    """CameraIntrin::depthToXYZ of the reference itself (Calibration.cpp compiled into oracle/_ref); intrin = (fx, cx, fy, cy);
    None if oracle/_ref is not built"""
    if not os.path.exists(REF_PAINTERS_PATH):
        return None
    lib = C.CDLL(REF_PAINTERS_PATH)
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    h, w = depth.shape
    out = np.zeros((h, w, 3), np.float32)
    lib.ref_depth_to_xyz.argtypes = [_P, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _P]
    lib.ref_depth_to_xyz(_p(depth), w, h, float(intrin[0]), float(intrin[1]), float(intrin[2]), float(intrin[3]), _p(out))
    return out


def paint_check_render(cloud, faces, vertex_part, width, height, intrin):
    """the product's rank-form renderer (avatar_b200/csrc/avb_paint.h) run on the CPU by tests/cpp/libpaint_check.so"""
    lib = C.CDLL(os.path.join(os.path.dirname(_HERE), "tests", "cpp", "libpaint_check.so"))
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vp = np.ascontiguousarray(vertex_part, dtype=np.uint8)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    depth = np.zeros((height, width), np.float32)
    parts = np.zeros((height, width), np.uint8)
    fids = np.zeros((height, width), np.int32)
    order = np.zeros(faces.shape[0], np.int32)
    lib.paint_check_render.argtypes = [_P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]
    lib.paint_check_render(_p(cloud), cloud.shape[0], _p(faces), faces.shape[0], _p(vp), width, height, _p(k), _p(depth),
                           _p(parts), _p(fids), _p(order))
    return dict(depth=depth, parts=parts, faces=fids, order=order)


def rtree_postprocess(image, roi, interval, num_parts, part_map_type=0, com_pre=None, dist_to_pre_weight=0.001):
    """RTree::postProcess restated (orc_rtree_postprocess): returns (image uint8 [H, W], com_pre [num_parts, 2]); com_pre=None
    is the reference's freshly resized matrix (x = -1, y = 0)"""
    img = np.array(image, dtype=np.uint8, order="C", copy=True)
    h, w = img.shape
    cp = np.tile(np.array([-1.0, 0.0]), (num_parts, 1)) if com_pre is None else np.array(com_pre, dtype=np.float64, order="C", copy=True)
    roi_a = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32)
    _lib.orc_rtree_postprocess.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_double]
    _lib.orc_rtree_postprocess.restype = None
    _lib.orc_rtree_postprocess(_p(img), w, h, _p(roi_a), int(interval), int(num_parts), int(part_map_type), _p(cp), float(dist_to_pre_weight))
    return img, cp


def brute_nn(points, queries):
    """exact brute-force 1-NN (nanoflann distance arithmetic, lowest index on exact ties)"""
    points, queries = _f64(points), _f64(queries)
    out = np.zeros(queries.shape[0], dtype=np.int32)
    _lib.orc_brute_nn.argtypes = [_P, C.c_int, _P, C.c_int, _P]
    _lib.orc_brute_nn.restype = None
    _lib.orc_brute_nn(_p(points), points.shape[0], _p(queries), queries.shape[0], _p(out))
    return out


REF_NANOFLANN_FMA_PATH = os.path.join(_HERE, "_ref", "libref_nanoflann_fma.so")


def paint_check_lambert(cloud, faces, width, height, intrin):
    """the product's rank-form renderLambert (avatar_b200/csrc/avb_paint.h) run on the CPU by tests/cpp/libpaint_check.so"""
    lib = C.CDLL(os.path.join(os.path.dirname(_HERE), "tests", "cpp", "libpaint_check.so"))
    cloud = _f64(cloud)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    k = np.ascontiguousarray(intrin, dtype=np.float32)
    gray = np.zeros((height, width), np.uint8)
    vlam = np.zeros(cloud.shape[0], np.float32)
    lib.paint_check_lambert.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]
    lib.paint_check_lambert.restype = None
    lib.paint_check_lambert(_p(cloud), cloud.shape[0], _p(faces), faces.shape[0], width, height, _p(k), _p(gray), _p(vlam))
    return gray, vlam


def ref_nanoflann_nn(points, queries, fma=False):
    """exact 1-NN through the reference's own vendored nanoflann (oracle/_ref); None if not built.
    fma=False: built with the reference's flags (CMakeLists.txt:37: -O3 -funroll-loops, no -march => no FMA);
    fma=True: the same source built with -march=x86-64-v3 (GCC contracts `result += diff*diff`), kept to measure how
    many answers depend on the contraction."""
    path = REF_NANOFLANN_FMA_PATH if fma else REF_NANOFLANN_PATH
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_nanoflann_nn.argtypes = [_P, C.c_int, _P, C.c_int, _P]
    points, queries = _f64(points), _f64(queries)
    out = np.zeros(queries.shape[0], dtype=np.int32)
    lib.ref_nanoflann_nn(_p(points), points.shape[0], _p(queries), queries.shape[0], _p(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# the reference's OWN Avatar.cpp + GaussianMixture.cpp (oracle/_ref/libref_avatar.so, compiled from /root/reference
# against the Eigen stand-in in oracle/shim): what the oracle's restatement of Avatar::update / GaussianMixture is
# pinned against
# ---------------------------------------------------------------------------------------------------------------
REF_AVATAR_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_avatar.so")
_ref_avatar = None


def _ref_avatar_lib():
    global _ref_avatar
    if _ref_avatar is None:
        if not os.path.exists(REF_AVATAR_PATH):
            return None
        L = C.CDLL(REF_AVATAR_PATH)
        L.ref_gmm_load.restype = _P
        L.ref_gmm_load.argtypes = [C.c_char_p]
        L.ref_gmm_free.argtypes = [_P]
        L.ref_gmm_dims.argtypes = [_P, _P, _P]
        L.ref_gmm_residual.restype = C.c_int
        L.ref_gmm_residual.argtypes = [_P, _P, _P]
        L.ref_gmm_pdf.restype = C.c_double
        L.ref_gmm_pdf.argtypes = [_P, _P]
        L.ref_gmm_tables.argtypes = [_P, _P, _P]
        L.ref_model_load.restype = _P
        L.ref_model_load.argtypes = [C.c_char_p]
        L.ref_model_dims.argtypes = [_P, _P]
        L.ref_model_tables.argtypes = [_P] * 11
        L.ref_model_free.argtypes = [_P]
        L.ref_avatar_update.argtypes = [_P] * 7
        L.ref_avatar_align.argtypes = [_P] * 6
        L.ref_opt_create.restype = _P
        L.ref_opt_create.argtypes = [_P, C.c_int, _P]
        L.ref_opt_free.argtypes = [_P]
        L.ref_opt_run.restype = C.c_int
        L.ref_opt_run.argtypes = [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                  C.c_int, _P, _P, _P, _P, _P]
        L.ref_render.argtypes = [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]
        _ref_avatar = L
    return _ref_avatar


def write_prior_text(path, weights, means, covs):
    """the reference's pose_prior.txt layout (GaussianMixture.cpp:12-44): nComps nDims, weights, means, covariances"""
    with open(path, "w") as fh:
        fh.write("%d %d\n" % means.shape)
        fh.write(" ".join(repr(float(v)) for v in weights) + "\n")
        for m in means:
            fh.write(" ".join(repr(float(v)) for v in m) + "\n")
        for c in covs:
            for row in c:
                fh.write(" ".join(repr(float(v)) for v in row) + "\n")


class RefGaussianMixture:
    """ark::GaussianMixture of the reference, loaded from a pose_prior.txt"""

    def __init__(self, path):
        L = _ref_avatar_lib()
        self.h = L.ref_gmm_load(path.encode())
        c, d = C.c_int(), C.c_int()
        L.ref_gmm_dims(self.h, C.byref(c), C.byref(d))
        self.C, self.D = c.value, d.value

    def residual(self, x):
        out = np.zeros(self.D + 1)
        comp = _ref_avatar_lib().ref_gmm_residual(self.h, _p(_f64(x)), _p(out))
        return out, comp

    def pdf(self, x):
        return _ref_avatar_lib().ref_gmm_pdf(self.h, _p(_f64(x)))

    def tables(self):
        pc, cl = np.zeros((self.C, self.D, self.D)), np.zeros(self.C)
        _ref_avatar_lib().ref_gmm_tables(self.h, _p(pc), _p(cl))
        return pc, cl


def write_model_dir(path, npz_path, prior_arrays):
    """data/avatar-model layout of the reference (AvatarModel.cpp:18-23): model.npz + pose_prior.txt"""
    import io
    import zipfile
    os.makedirs(path, exist_ok=True)
    # re-packed without the zip64 local headers numpy >= 1.x writes for every member: the reference's vendored cnpy
    # (cnpy.cpp) predates them and reads the 32-bit size fields only.  Same arrays, same dtypes, stored uncompressed.
    z = np.load(npz_path)
    with zipfile.ZipFile(os.path.join(path, "model.npz"), "w", zipfile.ZIP_STORED, allowZip64=False) as zf:
        for name in z.files:
            buf = io.BytesIO()
            np.lib.format.write_array(buf, np.ascontiguousarray(z[name]), version=(1, 0))
            zf.writestr(name + ".npy", buf.getvalue())
    write_prior_text(os.path.join(path, "pose_prior.txt"), prior_arrays["weights"], prior_arrays["means"], prior_arrays["covs"])
    return path


class RefAvatar:
    """ark::AvatarModel loaded by the reference's OWN AvatarModel.cpp (+ cnpy.cpp) from a model directory, and ark::Avatar
    over it"""

    def __init__(self, model_dir):
        L = _ref_avatar_lib()
        self.h = L.ref_model_load(model_dir.encode())
        d = np.zeros(8, np.int32)
        L.ref_model_dims(self.h, _p(d))
        self.V, self.J, self.K, self.F, self.C, self.D, self.use_jsr, self.n_assigned = (int(v) for v in d)

    def tables(self):
        """what the reference's loader derived: dict(parent, base, key, jsr_base, jsr, init_pos, mesh, asg_start, asg_joint, asg_weight)"""
        t = dict(parent=np.zeros(self.J, np.int32), base=np.zeros(3 * self.V), key=np.zeros((3 * self.V, self.K)),
                 jsr_base=np.zeros(3 * self.J), jsr=np.zeros((3 * self.J, self.K)), init_pos=np.zeros((self.J, 3)),
                 mesh=np.zeros((self.F, 3), np.int32), asg_start=np.zeros(self.V + 1, np.int32),
                 asg_joint=np.zeros(self.n_assigned, np.int32), asg_weight=np.zeros(self.n_assigned))
        _ref_avatar_lib().ref_model_tables(self.h, *(_p(t[k]) for k in ("parent", "base", "key", "jsr_base", "jsr", "init_pos", "mesh",
                                                                     "asg_start", "asg_joint", "asg_weight")))
        return t

    def update(self, p, R, w):
        cloud, jp, jt = np.zeros((self.V, 3)), np.zeros((self.J, 3)), np.zeros((self.J, 12))
        _ref_avatar_lib().ref_avatar_update(self.h, _p(_f64(p)), _p(_f64(R)), _p(_f64(w)), _p(cloud), _p(jp), _p(jt))
        return cloud, jp, jt

    def align_to_joints(self, pos):
        p, R, w0, smpl = np.zeros(3), np.zeros((self.J, 3, 3)), np.zeros(1), np.zeros(3 * (self.J - 1))
        _ref_avatar_lib().ref_avatar_align(self.h, _p(_f64(pos)), _p(p), _p(R), _p(w0), _p(smpl))
        return p, R, float(w0[0]), smpl


def ref_avatar_available():
    return _ref_avatar_lib() is not None


class RefOptimizer:
    """ark::AvatarOptimizer of the reference (AvatarOptimizer.cpp compiled from /root/reference): its own prologue, visibility,
    findNN over nanoflann, cost functors, evaluation callback and quaternion parameterization, over a model loaded by the
    reference's own AvatarModel.cpp.  The solver loop is the Levenberg-Marquardt restatement of oracle/ref_optimizer.cpp
    (Ceres is absent)."""

    def __init__(self, model_dir, num_parts, part_map):
        L = _ref_avatar_lib()
        self.avatar = RefAvatar(model_dir)      # the reference's own loader
        self.hm = self.avatar.h
        self.V, self.J, self.K, self.F = self.avatar.V, self.avatar.J, self.avatar.K, self.avatar.F
        self.part_map = np.ascontiguousarray(part_map, dtype=np.int32)
        self.h = L.ref_opt_create(self.hm, int(num_parts), _p(self.part_map))
        self.nx, self.P = 3 + 4 * self.J + self.K, 3 + 3 * self.J + self.K

    def _run(self, x, data, labels, icp_iters, max_iters, ftol, beta_pose, beta_shape, occlusion, threads, mode):
        x = np.array(x, dtype=np.float64, copy=True)
        data, labels = _f64(data), np.ascontiguousarray(labels, dtype=np.int32)
        cost, grad, H = np.zeros(1), np.zeros(self.P), np.zeros((self.P, self.P))
        stats, costs = np.zeros(3, np.int32), np.zeros(2)
        _ref_avatar_lib().ref_opt_run(self.h, _p(x), _p(data), _p(labels), data.shape[0], icp_iters, max_iters, ftol, beta_pose,
                                      beta_shape, int(occlusion), threads, mode, _p(cost), _p(grad), _p(H), _p(stats), _p(costs))
        return x, float(cost[0]), grad, H, stats, costs

    def render(self, cloud, width, height, intrin):
        """the reference's own AvatarRenderer::renderDepth / renderPartMask / renderFaces / renderLambert on a posed cloud"""
        cloud = _f64(cloud)
        k = np.ascontiguousarray(intrin, dtype=np.float32)
        out = dict(depth=np.zeros((height, width), np.float32), parts=np.zeros((height, width), np.uint8),
                   faces=np.zeros((height, width), np.int32), lambert=np.zeros((height, width), np.uint8))
        _ref_avatar_lib().ref_render(self.hm, _p(cloud), _p(k), width, height, _p(self.part_map), _p(out["depth"]), _p(out["parts"]),
                                     _p(out["faces"]), _p(out["lambert"]))
        return out

    def evaluate(self, x, data, labels, beta_pose=0.1, beta_shape=1.0, occlusion=True):
        """cost, gradient [P] and J^T J [P, P] of the reference's residual blocks at x (its own correspondences)"""
        _, cost, grad, H, stats, _ = self._run(x, data, labels, 1, 0, 0.0, beta_pose, beta_shape, occlusion, 1, 1)
        return cost, grad, H, int(stats[2])

    def optimize(self, x, data, labels, icp_iters=1, max_iters=10, function_tolerance=1e-4, beta_pose=0.1, beta_shape=1.0,
                 occlusion=True, threads=4):
        x, _, _, _, stats, costs = self._run(x, data, labels, icp_iters, max_iters, function_tolerance, beta_pose, beta_shape,
                                             occlusion, threads, 0)
        return x, {"iterations": int(stats[0]), "accepted": int(stats[1]), "initial_cost": float(costs[0]), "final_cost": float(costs[1])}
