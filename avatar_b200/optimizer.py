"""Avatar / AvatarOptimizer: host-side mirrors of ark::Avatar (include/Avatar.h:155-220) and
ark::AvatarOptimizer (include/AvatarOptimizer.h:11-54) over the C ABI.  Same member names, same
argument meaning, same in-place result semantics; all arithmetic runs in the CUDA library."""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import lib, check, ptr


def rotmat_to_quat(R):
    """AvatarOptimizer.cpp:1250-1254 (AngleAxisd::fromRotationMatrix -> Quaterniond), xyzw"""
    R = np.ascontiguousarray(R, dtype=np.float64)
    q = np.zeros(4)
    lib.avb_rotmat_to_quat(ptr(R), ptr(q))
    return q


def quat_to_rotmat(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    R = np.zeros((3, 3))
    lib.avb_quat_to_rotmat(ptr(q), ptr(R))
    return R


class Fitter:
    """Owns one avb_fitter (device buffers + stream) -- the device side of an AvatarOptimizer."""

    def __init__(self, model, num_parts, part_map, max_batch=1, max_total_points=1 << 18, device=0):
        self.model = model
        self.part_map = np.ascontiguousarray(part_map, dtype=np.int32)
        assert self.part_map.shape[0] == model.numJoints(), "partMap must have one entry per joint"
        cfg = _lib.FitterConfig(device, max_batch, max_total_points, num_parts, ptr(self.part_map))
        h = C.c_void_p()
        check(lib.avb_fitter_create(model.handle, C.byref(cfg), C.byref(h)))
        self.handle = h
        self.max_batch, self.max_total_points = max_batch, max_total_points
        self.nx = lib.avb_param_dim(model.handle)
        self.P = lib.avb_tangent_dim(model.handle)
        self.batch = 0
        self.total_points = 0

    def close(self):
        if getattr(self, "handle", None) is not None:
            lib.avb_fitter_destroy(self.handle)
            self.handle = None

    __del__ = close

    def avatar_update(self, x):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B, V, J = x.shape[0], self.model.numPoints(), self.model.numJoints()
        cloud, jp, jt = np.zeros((B, V, 3)), np.zeros((B, J, 3)), np.zeros((B, J, 12))
        check(lib.avb_avatar_update(self.handle, B, ptr(x), ptr(cloud), ptr(jp), ptr(jt)))
        return cloud, jp, jt

    def upload(self, clouds, labels, offsets):
        """float64 clouds go through avb_upload_batch; float32 clouds (what a depth camera delivers) through
        avb_upload_batch_f32: half the PCIe bytes, widened on the device to the same doubles"""
        f32 = isinstance(clouds, np.ndarray) and clouds.dtype == np.float32
        clouds = np.ascontiguousarray(clouds, dtype=np.float32 if f32 else np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._keep = (clouds, labels, offsets)
        self.batch = offsets.shape[0] - 1
        self.total_points = int(offsets[-1] - offsets[0])
        fn = lib.avb_upload_batch_f32 if f32 else lib.avb_upload_batch
        check(fn(self.handle, self.batch, ptr(clouds), ptr(labels), ptr(offsets)))

    def render(self, x, width, height, intrin, want=("depth", "parts", "faces")):
        """AvatarRenderer::renderDepth / renderPartMask / renderFaces of the model posed at x [B, nx]; intrin = (fx, cx, fy,
        cy).  Returns a dict of [B, H, W] images (float32 depth, uint8 parts, int32 faces = position in paint order)."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B = x.shape[0]
        d = _lib.RenderDesc(int(width), int(height), float(intrin[0]), float(intrin[1]), float(intrin[2]), float(intrin[3]))
        depth = np.zeros((B, height, width), np.float32) if "depth" in want else None
        parts = np.zeros((B, height, width), np.uint8) if "parts" in want else None
        faces = np.zeros((B, height, width), np.int32) if "faces" in want else None
        check(lib.avb_render_batch(self.handle, B, ptr(x), C.byref(d), ptr(depth), ptr(parts), ptr(faces)))
        return dict(depth=depth, parts=parts, faces=faces)

    def render_lambert(self, x, width, height, intrin):
        """AvatarRenderer::renderLambert of the model posed at x [B, nx]; intrin = (fx, cx, fy, cy) -> uint8 [B, H, W]"""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B = x.shape[0]
        d = _lib.RenderDesc(int(width), int(height), float(intrin[0]), float(intrin[1]), float(intrin[2]), float(intrin[3]))
        gray = np.zeros((B, height, width), np.uint8)
        check(lib.avb_render_lambert_batch(self.handle, B, ptr(x), C.byref(d), ptr(gray)))
        return gray

    def render_ms(self):
        ms = (C.c_float * 3)()
        check(lib.avb_last_render_ms(self.handle, ms))
        return dict(zip(("prepare", "cover", "resolve"), list(ms)))

    def set_rtree(self, tree, num_parts):
        """give the fitter a decision tree (RTree::nodes + leafBestMatch as arrays, see avb_rtree_desc)"""
        t = {k: np.ascontiguousarray(tree[k], dtype=d) for k, d in (("u", np.float32), ("v", np.float32), ("thresh", np.float32),
                                                                  ("lnode", np.int32), ("rnode", np.int32), ("leafid", np.int32),
                                                                  ("leaf_best", np.uint8))}
        d = _lib.RTreeDesc(len(t["thresh"]), len(t["leaf_best"]), int(num_parts), *(t[k].ctypes.data for k in
                           ("u", "v", "thresh", "lnode", "rnode", "leafid", "leaf_best")))
        check(lib.avb_fitter_set_rtree(self.handle, C.byref(d)))

    def rtree_predict(self, depth, roi=None, interval=1, fill_in_gaps=True):
        """RTree::predictBest on a batch of depth images [B,H,W] -> labels [B,H,W] uint8 (255 = not predicted)"""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        if depth.ndim == 2:
            depth = depth[None]
        B, H, W = depth.shape
        roi_a = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32).reshape(B, 4)
        out = np.zeros((B, H, W), dtype=np.uint8)
        check(lib.avb_rtree_predict_batch(self.handle, B, ptr(depth), W, H, ptr(roi_a), int(interval), int(bool(fill_in_gaps)),
                                          ptr(out)))
        return out

    def rtree_postprocess(self, parts, roi=None, interval=1, num_parts=16, part_map_type=0, com_pre=None, dist_to_pre_weight=0.001):
        """RTree::postProcess on a batch of label images [B,H,W] uint8 -> (images, com_pre [B, num_parts, 2])"""
        img = np.array(parts, dtype=np.uint8, order="C", copy=True)
        if img.ndim == 2:
            img = img[None]
        B, H, W = img.shape
        cp = (np.tile(np.array([-1.0, 0.0]), (B, num_parts, 1)) if com_pre is None
              else np.array(com_pre, dtype=np.float64, order="C", copy=True).reshape(B, num_parts, 2))
        roi_a = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32).reshape(B, 4)
        check(lib.avb_rtree_postprocess_batch(self.handle, B, ptr(img), W, H, ptr(roi_a), int(interval), int(num_parts),
                                              int(part_map_type), ptr(cp), float(dist_to_pre_weight)))
        return img, cp

    def rtree_ms(self):
        ms = C.c_float()
        check(lib.avb_last_rtree_ms(self.handle, C.byref(ms)))
        return ms.value

    def upload_depth(self, depth, parts, intrin, num_parts, roi=None, interval=1, rtree_interval=2, postprocess=False,
                     part_map_type=0, dist_to_pre_weight=0.001):
        """Build the batch's data clouds on the device from depth [B,H,W] float32 (metres) and body-part label images
        [B,H,W] uint8 (255 = background), as demo.cpp:215-250 + CameraIntrin::depthToXYZ do on the host.
        intrin = (fx, cx, fy, cy); roi: optional [B,4] int32 (x0, y0, x1, y1 inclusive).  Returns the offsets."""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        if depth.ndim == 2:
            depth = depth[None]
        if parts is not None:   # parts=None: labels from the fitter's decision tree (set_rtree), on the device
            parts = np.ascontiguousarray(parts, dtype=np.uint8).reshape(depth.shape)
        B, H, W = depth.shape
        roi_a = None if roi is None else np.ascontiguousarray(roi, dtype=np.int32).reshape(B, 4)
        img = _lib.ImageDesc(W, H, float(intrin[0]), float(intrin[1]), float(intrin[2]), float(intrin[3]), int(interval),
                             int(num_parts), int(rtree_interval), int(bool(postprocess)), int(part_map_type), float(dist_to_pre_weight))
        off = np.zeros(B + 1, dtype=np.int64)
        self._keep = (depth, parts, roi_a)
        check(lib.avb_upload_depth_batch(self.handle, B, ptr(depth), ptr(parts), ptr(roi_a), C.byref(img), ptr(off)))
        self.batch = B
        self.total_points = int(off[-1])
        return off

    def cloud_ms(self):
        """device ms of (cloud_count_kernel, cloud_compact_kernel) of the last upload_depth"""
        ms = (C.c_float * 2)()
        check(lib.avb_last_cloud_ms(self.handle, ms))
        return float(ms[0]), float(ms[1])

    def download_batch(self):
        """the resident batch: (clouds [N,3] float64, labels [N] int32, offsets [B+1])"""
        pts = np.zeros((self.total_points, 3))
        lab = np.zeros(self.total_points, dtype=np.int32)
        off = np.zeros(self.batch + 1, dtype=np.int64)
        check(lib.avb_download_batch(self.handle, ptr(pts), ptr(lab), ptr(off)))
        return pts, lab, off

    def fit_resident(self, x, opt):
        x = np.ascontiguousarray(x, dtype=np.float64)
        check(lib.avb_fit_resident(self.handle, ptr(x), C.byref(opt)))

    def download(self, want_cloud=False):
        x = np.zeros((self.batch, self.nx))
        stats = (_lib.Stats * self.batch)()
        cloud = np.zeros((self.batch, self.model.numPoints(), 3)) if want_cloud else None
        check(lib.avb_download_results(self.handle, ptr(x), stats, ptr(cloud)))
        return x, list(stats), cloud

    def fit_batch(self, clouds, labels, offsets, x, opt, want_cloud=False):
        clouds = np.ascontiguousarray(clouds, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        x = np.array(x, dtype=np.float64, order="C", copy=True).reshape(offsets.shape[0] - 1, self.nx)
        B = x.shape[0]
        stats = (_lib.Stats * B)()
        cloud = np.zeros((B, self.model.numPoints(), 3)) if want_cloud else None
        check(lib.avb_fit_batch(self.handle, B, ptr(clouds), ptr(labels), ptr(offsets), ptr(x), C.byref(opt), stats,
                                ptr(cloud)))
        self.batch = B
        self.total_points = int(offsets[-1] - offsets[0])
        return x, list(stats), cloud

    def track_sequence(self, clouds, labels, offsets, x0, opt):
        """configs[3]: fit T frames in order, each warm-started from the previous fit; returns x [T][nx], stats"""
        clouds = np.ascontiguousarray(clouds, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        T = offsets.shape[0] - 1
        x = np.zeros((T, self.nx))
        stats = (_lib.Stats * T)()
        check(lib.avb_track_sequence(self.handle, T, ptr(clouds), ptr(labels), ptr(offsets), ptr(x0), C.byref(opt),
                                     ptr(x), stats))
        self.batch = 1
        self.total_points = int(offsets[-1] - offsets[-2])
        return x, list(stats)

    # parity taps
    def debug_correspond(self, x, opt):
        x = np.ascontiguousarray(x, dtype=np.float64)
        check(lib.avb_debug_correspond(self.handle, ptr(x), C.byref(opt)))

    def debug_read(self, what):
        B, V = self.batch, self.model.numPoints()
        shape, dt = {_lib.TAP_VISIBLE: ((B, V), np.uint8), _lib.TAP_NN: ((self.total_points,), np.int32),
                     _lib.TAP_CLOUD: ((B, V, 3), np.float64), _lib.TAP_COUNT: ((B, V), np.int32),
                     _lib.TAP_SUM: ((B, V, 3), np.float64)}[what]
        out = np.zeros(shape, dtype=dt)
        check(lib.avb_debug_read(self.handle, what, ptr(out), out.nbytes))
        return out

    def debug_evaluate(self, x, opt):
        x = np.ascontiguousarray(x, dtype=np.float64)
        B, P = self.batch, self.P
        cost, grad, H = np.zeros(B), np.zeros((B, P)), np.zeros((B, P, P))
        check(lib.avb_debug_evaluate(self.handle, ptr(x), C.byref(opt), ptr(cost), ptr(grad), ptr(H)))
        return cost, grad, H

    def device_ms(self):
        tot = C.c_float()
        per = (C.c_float * 4)()
        check(lib.avb_last_device_ms(self.handle, C.byref(tot), per))
        return tot.value, list(per)

    KERNEL_CLASSES = ("pose_visibility_kernel", "nn_kernel", "lm_prep_kernel", "lm_rows_kernel", "lm_gram_kernel",
                      "lm_solve_kernel", "pose_visibility_kernel(final)", "lm_flow_kernel")

    def set_profiling(self, on):
        check(lib.avb_set_profiling(self.handle, int(bool(on))))

    def kernel_ms(self):
        """{kernel class: (total ms, launches)} of the last profiled fit_resident"""
        ms = (C.c_float * 8)()
        n = (C.c_int32 * 8)()
        check(lib.avb_last_kernel_ms(self.handle, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.KERNEL_CLASSES) if n[i] > 0}

    def groups(self):
        """[(joints, model vertices)] of the static Jacobian column groups"""
        n = C.c_int32()
        nj = (C.c_int32 * 16)()
        nv = (C.c_int32 * 16)()
        check(lib.avb_fitter_groups(self.handle, C.byref(n), nj, nv))
        return [(nj[g], nv[g]) for g in range(n.value)]

    def flow_task_ms(self):
        """CTA milliseconds the last profiled lm_flow_kernel spent in (record tasks, Gram tasks, solves, waiting)"""
        ms = (C.c_float * 4)()
        check(lib.avb_last_flow_task_ms(self.handle, ms))
        return dict(zip(("rows", "gram", "solve", "wait"), list(ms)))

    def flow_phase_ms(self):
        """finer split of flow_task_ms (CTA ms per sub-phase of the solve / record / Gram tasks)"""
        ms = (C.c_float * 12)()
        check(lib.avb_last_flow_phase_ms(self.handle, ms))
        names = ("solve.load", "solve.reduce", "solve.basis", "solve.prior", "solve.control", "solve.cholesky",
                 "solve.backsub", "solve.retract", "rows.prologue", "rows.body", "gram.load", "gram.dmma")
        return dict(zip(names, list(ms)))

    def timer_start(self):
        check(lib.avb_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        check(lib.avb_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def comm_init(self, unique_id, rank, nranks):
        """join the NCCL communicator described by the 128-byte id of avatar_b200.shard.comm_unique_id()"""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(lib.avb_fitter_comm_init(self.handle, buf, int(rank), int(nranks)))
        self.nranks = int(nranks)

    def gather_params(self):
        """ncclAllGather of the device parameter blocks of the last fit -> [nranks, max_batch, nx] (host)"""
        out = np.zeros((self.nranks, self.max_batch, self.nx))
        check(lib.avb_gather_params(self.handle, ptr(out)))
        return out

    def gather_begin(self):
        """first half of gather_params: snapshot + collective on its own stream, returns at once"""
        check(lib.avb_gather_params_begin(self.handle))

    def gather_end(self):
        out = np.zeros((self.nranks, self.max_batch, self.nx))
        check(lib.avb_gather_params_end(self.handle, ptr(out)))
        return out

    def synchronize(self):
        check(lib.avb_synchronize(self.handle))

    def launch_count(self):
        return lib.avb_last_launch_count(self.handle)


class Avatar:
    """ark::Avatar: p, r (rotation matrices), w in; cloud, jointPos, jointTrans out (update())."""

    def __init__(self, model):
        self.model = model
        self.w = np.zeros(model.numShapeKeys())
        self.p = np.zeros(3)
        self.r = [np.eye(3) for _ in range(model.numJoints())]
        self.cloud = np.zeros((3, 0))
        self.jointPos = np.zeros((3, 0))
        self.jointTrans = np.zeros((12, 0))
        self._fitter = None

    def _updater(self):
        if self._fitter is None:
            self._fitter = Fitter(self.model, 1, np.zeros(self.model.numJoints(), dtype=np.int32), 1, 16)
        return self._fitter

    def params(self):
        """x = [p | q (xyzw per joint, via the reference's R->AngleAxis->Quaternion prologue) | w]"""
        q = np.concatenate([rotmat_to_quat(R) for R in self.r])
        return np.concatenate([self.p, q, self.w])

    def set_params(self, x):
        J = self.model.numJoints()
        self.p = np.array(x[:3])
        self.r = [quat_to_rotmat(x[3 + 4 * j:7 + 4 * j]) for j in range(J)]  # AvatarOptimizer.cpp:1494-1496
        self.w = np.array(x[3 + 4 * J:])

    def update(self, fitter=None):
        """Avatar::update (Avatar.cpp:22-75) on the GPU"""
        cloud, jp, jt = (fitter or self._updater()).avatar_update(self.params())
        self.cloud, self.jointPos, self.jointTrans = cloud[0].T.copy(), jp[0].T.copy(), jt[0].T.copy()

    def smplParams(self):
        """Avatar::smplParams (Avatar.cpp:128-137): axis-angle of joints 1..J-1"""
        out = []
        for R in self.r[1:]:
            q = rotmat_to_quat(R)
            n = np.linalg.norm(q[:3])
            out.append(np.zeros(3) if n == 0 else q[:3] / n * (2 * np.arctan2(n, abs(q[3]))))
        return np.concatenate(out)


class AvatarOptimizer:
    """ark::AvatarOptimizer (include/AvatarOptimizer.h): optimize(data_cloud, data_part_labels, icp_iters, num_threads)"""
    ROT_SIZE = 4

    def __init__(self, ava, intrin, image_size, num_parts, part_map, max_points=1 << 18, device=0):
        self.ava, self.intrin, self.imageSize = ava, intrin, image_size
        self.numParts, self.partMap = num_parts, part_map
        self.betaPose, self.betaShape = 0.1, 1.0     # AvatarOptimizer.h:27
        self.nnStep = 20                              # :33
        self.maxItersPerICP = 10                      # :36
        self.enableOcclusion = True                   # :39
        self.functionTolerance = 1e-4                 # AvatarOptimizer.cpp:1333
        self.r = [np.array([0.0, 0.0, 0.0, 1.0]) for _ in range(ava.model.numJoints())]
        self.fitter = Fitter(ava.model, num_parts, part_map, 1, max_points, device)
        self.last_stats = None

    def options(self, icp_iters):
        o = _lib.default_options()
        o.icp_iters, o.max_iters_per_icp = icp_iters, self.maxItersPerICP
        o.beta_pose, o.beta_shape = self.betaPose, self.betaShape
        o.enable_occlusion, o.nn_step = int(self.enableOcclusion), self.nnStep
        o.function_tolerance = self.functionTolerance
        return o

    def optimize(self, data_cloud, data_part_labels, icp_iters=1, num_threads=4):
        """data_cloud: 3 x N (column = point), as Eigen::Matrix<double,3,Dynamic>; labels: N ints.
        num_threads is accepted for source compatibility and unused (the GPU does the work)."""
        data_cloud = np.asarray(data_cloud, dtype=np.float64)
        assert data_cloud.shape[0] == 3
        pts = np.ascontiguousarray(data_cloud.T)       # column-major 3xN == row-major Nx3
        labels = np.ascontiguousarray(data_part_labels, dtype=np.int32)
        x = self.ava.params()                          # prologue :1250-1254
        N = pts.shape[0]
        xs, stats, cloud = self.fitter.fit_batch(pts, labels, np.array([0, N]), x[None], self.options(icp_iters), True)
        J = self.ava.model.numJoints()
        self.r = [xs[0, 3 + 4 * j:7 + 4 * j].copy() for j in range(J)]
        self.ava.set_params(xs[0])                     # :1494-1496
        self.ava.update(self.fitter)                   # :1497
        self.last_stats = stats[0]
