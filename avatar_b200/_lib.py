"""ctypes binding of the C ABI in include/avatar_b200.h (libavatar_b200.so, built in-tree).

The library is the product: there is no Python or CPU fallback.  Import fails loudly when the
shared object is missing, and every compute call fails with AvbError when no B200 is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavatar_b200.so")


class AvbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"avatar_b200 error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("num_points", C.c_int32), ("num_joints", C.c_int32), ("num_shape_keys", C.c_int32),
                ("num_faces", C.c_int32),
                ("base_cloud", C.c_void_p), ("key_clouds", C.c_void_p), ("joint_shape_reg_base", C.c_void_p),
                ("joint_shape_reg", C.c_void_p), ("parent", C.c_void_p), ("mesh", C.c_void_p),
                ("assign_start", C.c_void_p), ("assign_joint", C.c_void_p), ("assign_weight", C.c_void_p),
                ("gmm_components", C.c_int32), ("gmm_dims", C.c_int32),
                ("gmm_weight", C.c_void_p), ("gmm_mean", C.c_void_p), ("gmm_cov", C.c_void_p)]


class FitterConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_batch", C.c_int32), ("max_total_points", C.c_int64),
                ("num_parts", C.c_int32), ("part_map", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("icp_iters", C.c_int32), ("max_iters_per_icp", C.c_int32), ("beta_pose", C.c_double),
                ("beta_shape", C.c_double), ("enable_occlusion", C.c_int32), ("nn_step", C.c_int32),
                ("function_tolerance", C.c_double), ("solver", C.c_int32), ("jtj_precision", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class Stats(C.Structure):
    _fields_ = [("num_points", C.c_int32), ("num_correspondences", C.c_int32),
                ("num_matched_vertices", C.c_int32), ("iterations", C.c_int32),
                ("accepted_steps", C.c_int32), ("status", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]


# every symbol include/avatar_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
class ImageDesc(C.Structure):   # avb_image_desc
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("cx", C.c_float), ("fy", C.c_float),
                ("cy", C.c_float), ("interval", C.c_int32), ("num_parts", C.c_int32), ("rtree_interval", C.c_int32),
                ("rtree_postprocess", C.c_int32), ("part_map_type", C.c_int32), ("dist_to_pre_weight", C.c_double)]


class RenderDesc(C.Structure):   # avb_render_desc
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("cx", C.c_float), ("fy", C.c_float),
                ("cy", C.c_float)]


class RTreeDesc(C.Structure):   # avb_rtree_desc
    _fields_ = [("num_nodes", C.c_int32), ("num_leaves", C.c_int32), ("num_parts", C.c_int32), ("u", C.c_void_p),
                ("v", C.c_void_p), ("thresh", C.c_void_p), ("lnode", C.c_void_p), ("rnode", C.c_void_p),
                ("leafid", C.c_void_p), ("leaf_best", C.c_void_p)]


SYMBOLS = [
    ("avb_default_options", None, [C.POINTER(Options)]),
    ("avb_last_error", C.c_char_p, []),
    ("avb_device_count", C.c_int, []),
    ("avb_model_create", C.c_int, [C.POINTER(ModelDesc), C.POINTER(_P)]),
    ("avb_model_destroy", None, [_P]),
    ("avb_model_dims", C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("avb_model_get_prior", C.c_int, [_P, _P, _P]),
    ("avb_fitter_create", C.c_int, [_P, C.POINTER(FitterConfig), C.POINTER(_P)]),
    ("avb_fitter_destroy", None, [_P]),
    ("avb_param_dim", C.c_int, [_P]),
    ("avb_tangent_dim", C.c_int, [_P]),
    ("avb_rotmat_to_quat", None, [_P, _P]),
    ("avb_quat_to_rotmat", None, [_P, _P]),
    ("avb_avatar_update", C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    ("avb_fit", C.c_int, [_P, _P, _P, C.c_int32, _P, C.POINTER(Options), _P, _P]),
    ("avb_fit_batch", C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.POINTER(Options), _P, _P]),
    ("avb_track_sequence", C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.POINTER(Options), _P, _P]),
    ("avb_upload_batch", C.c_int, [_P, C.c_int32, _P, _P, _P]),
    ("avb_upload_batch_f32", C.c_int, [_P, C.c_int32, _P, _P, _P]),
    ("avb_fit_resident", C.c_int, [_P, _P, C.POINTER(Options)]),
    ("avb_upload_depth_batch", C.c_int, [_P, C.c_int32, _P, _P, _P, C.POINTER(ImageDesc), _P]),
    ("avb_download_batch", C.c_int, [_P, _P, _P, _P]),
    ("avb_last_cloud_ms", C.c_int, [_P, _P]),
    ("avb_fitter_set_rtree", C.c_int, [_P, C.POINTER(RTreeDesc)]),
    ("avb_rtree_predict_batch", C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    ("avb_last_rtree_ms", C.c_int, [_P, _P]),
    ("avb_rtree_postprocess_batch", C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_double]),
    ("avb_rtree_reset_tracking", C.c_int, [_P]),
    ("avb_render_batch", C.c_int, [_P, C.c_int32, _P, C.POINTER(RenderDesc), _P, _P, _P]),
    ("avb_last_render_ms", C.c_int, [_P, _P]),
    ("avb_render_lambert_batch", C.c_int, [_P, C.c_int32, _P, C.POINTER(RenderDesc), _P]),
    ("avb_download_results", C.c_int, [_P, _P, _P, _P]),
    ("avb_synchronize", C.c_int, [_P]),
    ("avb_comm_unique_id", C.c_int, [_P]),
    ("avb_fitter_comm_init", C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    ("avb_gather_params", C.c_int, [_P, _P]),
    ("avb_gather_params_begin", C.c_int, [_P]),
    ("avb_gather_params_end", C.c_int, [_P, _P]),
    ("avb_last_device_ms", C.c_int, [_P, C.POINTER(C.c_float), _P]),
    ("avb_timer_start", C.c_int, [_P]),
    ("avb_timer_stop", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("avb_set_profiling", C.c_int, [_P, C.c_int]),
    ("avb_last_kernel_ms", C.c_int, [_P, _P, _P]),
    ("avb_last_flow_task_ms", C.c_int, [_P, _P]),
    ("avb_last_flow_phase_ms", C.c_int, [_P, _P]),
    ("avb_fitter_groups", C.c_int, [_P, _P, _P, _P]),
    ("avb_last_launch_count", C.c_int, [_P]),
    ("avb_host_alloc", _P, [C.c_uint64]),
    ("avb_host_free", None, [_P]),
    ("avb_debug_correspond", C.c_int, [_P, _P, C.POINTER(Options)]),
    ("avb_debug_read", C.c_int, [_P, C.c_int, _P, C.c_uint64]),
    ("avb_debug_evaluate", C.c_int, [_P, _P, C.POINTER(Options), _P, _P, _P]),
]

JTJ_FP64, JTJ_FP32, JTJ_BF16_TENSOR = 0, 1, 2
TAP_VISIBLE, TAP_NN, TAP_CLOUD, TAP_COUNT, TAP_SUM = 1, 2, 3, 4, 5

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or make -C avatar_b200/csrc). avatar_b200 has no fallback implementation.")
lib = C.CDLL(LIB_PATH)
for _name, _res, _args in SYMBOLS:
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc):
    if rc != 0:
        raise AvbError(rc, (lib.avb_last_error() or b"").decode())


def ptr(a):
    """raw pointer of a C-contiguous numpy array (or None)"""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_options():
    o = Options()
    lib.avb_default_options(C.byref(o))
    return o
