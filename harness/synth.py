"""Synthetic smplsynth-style frames: the recipe of Avatar::randomize
(Avatar.cpp:77-126), the optim.cpp:74-145 scenario (render -> back-project -> perturbed start) and
the fixed 640x576 intrinsics.

TEST / BENCH HARNESS, not product: the rasteriser is host C++ in harness/libavb_harness.so (its own shared object, so
that a process that only generates fixtures -- bench.py --impl reference -- never maps the product library), and
nothing here imports avatar_b200.  Functions take any model object with the reference's member names
(numJoints(), numShapeKeys(), posePrior, assignedJoints, mesh): avatar_b200.AvatarModel or the plain-numpy HostModel
below."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavb_harness.so")
if not os.path.exists(LIB_PATH):
    subprocess.check_call(["make", "-s", "-C", _HERE])
lib = C.CDLL(LIB_PATH)
_P = C.c_void_p
lib.avb_synth_render.restype = C.c_int
lib.avb_synth_render.argtypes = [_P, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float,
                                 C.c_float, _P, _P]
lib.avb_synth_backproject.restype = C.c_int64
lib.avb_synth_backproject.argtypes = [_P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                      _P, _P, C.c_int64]


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(rc):
    if rc != 0:
        raise RuntimeError(f"harness call failed with code {rc}")


class _Prior:
    pass


class HostModel:
    """the members of ark::AvatarModel the generators read, loaded with numpy only (model.npz schema of
    AvatarModel.cpp:26-102); used where the product library must not be loaded"""

    def __init__(self, npz_path, prior_arrays):
        npz = np.load(npz_path)
        weights = np.asarray(npz["weights"], dtype=np.float64)
        self.parent = np.asarray(npz["kintree_table"])[0].astype(np.uint32).astype(np.int32)
        self.mesh = np.ascontiguousarray(np.asarray(npz["f"]).astype(np.int64).astype(np.int32))
        self._K = int(np.asarray(npz["shapedirs"]).shape[2])
        self.assignedJoints = []
        for v in range(weights.shape[0]):   # (weight, joint), weight > 1e-12, descending by the pair (AvatarModel.cpp:74-94)
            nz = np.nonzero(weights[v] > 1e-12)[0]
            self.assignedJoints.append(sorted(((float(weights[v, j]), int(j)) for j in nz), reverse=True))
        g = _Prior()
        g.weight = np.ascontiguousarray(prior_arrays["weights"], dtype=np.float64)
        g.mean = np.ascontiguousarray(prior_arrays["means"], dtype=np.float64)
        g.cov = np.ascontiguousarray(prior_arrays["covs"], dtype=np.float64)
        g.nComps, g.nDims = g.mean.shape
        self.posePrior = g

    def numJoints(self): return int(self.parent.shape[0])
    def numShapeKeys(self): return self._K

WIDTH, HEIGHT = 640, 576
FX = FY = 504.0
CX, CY = 320.0, 288.0


def _aa_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    return np.concatenate([np.sin(0.5 * angle) * axis, [np.cos(0.5 * angle)]])


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _from_spherical(rho, theta, phi):  # AvatarHelpers.cpp fromSpherical
    return np.array([rho * np.sin(phi) * np.cos(theta), rho * np.cos(phi), rho * np.sin(phi) * np.sin(theta)])


def random_params(model, rng, pose_scale=1.0):
    """GT parameters x = [p | q | w] following Avatar::randomize (Avatar.cpp:84-125)."""
    J, K = model.numJoints(), model.numShapeKeys()
    w = rng.standard_normal(K)
    g = model.posePrior
    comp = rng.choice(g.nComps, p=g.weight / g.weight.sum())
    L = np.linalg.cholesky(g.cov[comp])
    samp = g.mean[comp] + pose_scale * (L @ rng.standard_normal(g.nDims))
    q = [None] * J
    for i in range(J - 1):
        a = samp[3 * i:3 * i + 3]
        ang = np.linalg.norm(a)
        q[i + 1] = _aa_quat(a / ang, ang) if ang > 0 else np.array([0, 0, 0, 1.0])
    p = np.array([rng.uniform(-1.0, 1.0), rng.uniform(-0.5, 0.5), rng.uniform(2.2, 4.5)])
    angle_up = rng.uniform(-np.pi / 3, np.pi / 3) + np.pi
    q_up = _aa_quat([0, 1, 0], angle_up)
    axis_p = _from_spherical(1.0, rng.uniform(0, 2 * np.pi), rng.uniform(-np.pi / 2, np.pi / 2))
    q_pert = _aa_quat(axis_p, rng.normal(0.0, 0.2))
    q[0] = _qmul(q_pert, q_up)
    return np.concatenate([p, np.concatenate(q), w])


def perturbed_start(model, x_gt, rng, rot_sigma=0.1):
    """optim.cpp:122-145: every rotation right-multiplied by a N(0, 0.1 rad) twist, w = 0 except w[0] = -2.5"""
    J, K = model.numJoints(), model.numShapeKeys()
    x = x_gt.copy()
    for j in range(J):
        axis = _from_spherical(1.0, rng.uniform(0, 2 * np.pi), rng.uniform(-np.pi / 2, np.pi / 2))
        dq = _aa_quat(axis, rng.normal(0.0, rot_sigma))
        q = _qmul(x[3 + 4 * j:7 + 4 * j], dq)
        if q[3] < 0:
            q = -q   # the optimize() prologue always starts from w >= 0
        x[3 + 4 * j:7 + 4 * j] = q
    x[3 + 4 * J:] = 0.0
    x[3 + 4 * J] = -2.5
    return x


def slerp_sequence(model, rng, frames, key_every=30, pose_scale=0.6):
    """BASELINE.json configs[3]: ground-truth parameters of a smooth motion -- random key poses (Avatar::randomize style) every
    `key_every` frames, root translation interpolated linearly, every joint rotation by spherical linear interpolation,
    shape constant.  Returns [frames][nx]."""
    J = model.numJoints()
    nkeys = (frames + key_every - 1) // key_every + 1
    keys = [random_params(model, rng, pose_scale=pose_scale) for _ in range(nkeys)]
    for k in range(1, nkeys):                       # same shape, nearby root, same hemisphere for every quaternion
        keys[k][3 + 4 * J:] = keys[0][3 + 4 * J:]
        keys[k][:3] = keys[k - 1][:3] + rng.uniform(-0.15, 0.15, 3)
        q_up = _aa_quat([0, 1, 0], rng.normal(0.0, 0.3))
        keys[k][3:7] = _qmul(q_up, keys[k - 1][3:7])
        for j in range(J):
            a, b = keys[k - 1][3 + 4 * j:7 + 4 * j], keys[k][3 + 4 * j:7 + 4 * j]
            if np.dot(a, b) < 0:
                keys[k][3 + 4 * j:7 + 4 * j] = -b
    out = np.zeros((frames, keys[0].shape[0]))
    for t in range(frames):
        k, u = divmod(t, key_every)
        u = u / key_every
        a, b = keys[k], keys[k + 1]
        x = (1 - u) * a + u * b                       # translation and shape
        for j in range(J):
            qa, qb = a[3 + 4 * j:7 + 4 * j], b[3 + 4 * j:7 + 4 * j]
            d = float(np.clip(np.dot(qa, qb), -1.0, 1.0))
            th = np.arccos(d)
            q = qa if th < 1e-9 else (np.sin((1 - u) * th) * qa + np.sin(u * th) * qb) / np.sin(th)
            x[3 + 4 * j:7 + 4 * j] = q / np.linalg.norm(q)
        out[t] = x
    return out


def vertex_parts(model, part_map):
    return np.array([part_map[pairs[0][1]] for pairs in model.assignedJoints], dtype=np.int32)


def render_cloud(model, cloud_V3, part_map, width=WIDTH, height=HEIGHT, fx=FX, fy=FY, cx=CX, cy=CY, interval=1):
    """z-buffer render of a posed mesh, then back-projection of every `interval`-th hit pixel."""
    cloud = np.ascontiguousarray(cloud_V3, dtype=np.float64)
    vp = vertex_parts(model, part_map)
    faces = np.ascontiguousarray(model.mesh, dtype=np.int32)
    depth = np.zeros((height, width), dtype=np.float32)
    part = np.zeros((height, width), dtype=np.uint8)
    check(lib.avb_synth_render(ptr(cloud), cloud.shape[0], ptr(faces), faces.shape[0], ptr(vp), width, height,
                               fx, fy, cx, cy, ptr(depth), ptr(part)))
    return backproject(depth, part, fx, fy, cx, cy, interval) + (depth, part)


def backproject(depth, part, fx=FX, fy=FY, cx=CX, cy=CY, interval=1):
    h, w = depth.shape
    cap = int((depth > 0).sum()) + 1
    pts = np.zeros((cap, 3))
    lab = np.zeros(cap, dtype=np.int32)
    n = lib.avb_synth_backproject(ptr(depth), ptr(part), w, h, fx, fy, cx, cy, interval, ptr(pts), ptr(lab), cap)
    assert n >= 0
    return pts[:n].copy(), lab[:n].copy()


def random_rtree(rng, num_parts, depth_levels=12, leaf_prob=0.12):
    """a random decision tree in the layout of RTree::nodes / leafBestMatch (include/RTree.h:28-41): probe offsets in
    pixel-metres (a few to ~150 px at 1 m), thresholds around zero, children after their parent.  Harness data: the
    reference ships no trained tree."""
    u, v, thresh, lnode, rnode, leafid, leaf_best = [], [], [], [], [], [], []

    def new_node():
        u.append([0.0, 0.0]); v.append([0.0, 0.0]); thresh.append(0.0); lnode.append(-1); rnode.append(-1); leafid.append(-1)
        return len(thresh) - 1
    frontier = [(new_node(), 0)]
    while frontier:
        nxt = []
        for n, d in frontier:
            if d >= depth_levels or (d >= 3 and rng.random() < leaf_prob):
                leafid[n] = len(leaf_best)
                leaf_best.append(int(rng.integers(0, num_parts)))
                continue
            u[n] = list(rng.normal(0.0, 60.0, 2))
            v[n] = list(rng.normal(0.0, 60.0, 2)) if rng.random() < 0.8 else [0.0, 0.0]
            thresh[n] = float(rng.normal(0.0, 0.4)) if rng.random() < 0.7 else float(rng.normal(0.0, 8.0))
            lnode[n], rnode[n] = new_node(), new_node()
            nxt += [(lnode[n], d + 1), (rnode[n], d + 1)]
        frontier = nxt
    return dict(u=np.array(u, np.float32), v=np.array(v, np.float32), thresh=np.array(thresh, np.float32),
                lnode=np.array(lnode, np.int32), rnode=np.array(rnode, np.int32), leafid=np.array(leafid, np.int32),
                leaf_best=np.array(leaf_best, np.uint8))
