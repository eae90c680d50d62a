"""avatar_b200: B200-native SMPL-to-point-cloud fitting engine behind the sxyu/avatar
Avatar / AvatarModel / AvatarOptimizer interface.  The compute lives in libavatar_b200.so
(hand-written sm_100a CUDA + a C ABI, include/avatar_b200.h); these modules only mirror the
reference's host-side interface.  No CPU fallback exists."""
from ._lib import AvbError, default_options  # noqa: F401  (import fails loudly if the .so is missing)
from .gmm import GaussianMixture  # noqa: F401
from .model import AvatarModel  # noqa: F401
from .optimizer import Avatar, AvatarOptimizer, Fitter, rotmat_to_quat, quat_to_rotmat  # noqa: F401
