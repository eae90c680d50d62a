"""AvatarModel: host-side mirror of ark::AvatarModel (include/Avatar.h:64-151).

Restates the npz branch of the reference constructor (AvatarModel.cpp:23-127): same member names,
same layouts, same derived quantities (assignedJoints, jointShapeRegBase, jointShapeReg).  The
arrays are handed to the library through avb_model_create; nothing is computed on the CPU at fit
time.
"""
import ctypes as C
import os
import numpy as np

from . import _lib
from .gmm import GaussianMixture


class AvatarModel:
    def __init__(self, model_dir="", limit_one_joint_per_point=False, npz_path=None, pose_prior=None):
        """model_dir: directory holding model.npz (+ pose_prior.txt), as in the reference.
        limit_one_joint_per_point is accepted and ignored on the npz path, like the reference
        (AvatarModel.cpp:23-127 never reads it)."""
        self.MODEL_DIR = model_dir
        if npz_path is None:
            root = model_dir or os.path.join(os.environ.get("OPENARK_DIR", "."), "data", "avatar-model")
            npz_path = os.path.join(root, "model.npz")
            prior_path = os.path.join(root, "pose_prior.txt")
        else:
            prior_path = None
        npz = np.load(npz_path)
        v_template = np.asarray(npz["v_template"], dtype=np.float64)            # (V,3)
        shapedirs = np.asarray(npz["shapedirs"], dtype=np.float64)              # (V,3,K)
        faces = np.asarray(npz["f"]).astype(np.int64).astype(np.int32)           # (F,3)
        kintree = np.asarray(npz["kintree_table"])
        jreg = np.asarray(npz["J_regressor"], dtype=np.float64)                 # (J,V)
        weights = np.asarray(npz["weights"], dtype=np.float64)                  # (V,J)
        V, J, K, F = v_template.shape[0], kintree.shape[1], shapedirs.shape[2], faces.shape[0]
        assert v_template.shape == (V, 3) and shapedirs.shape == (V, 3, K)
        assert jreg.shape == (J, V) and weights.shape == (V, J) and faces.shape == (F, 3)
        # kintree row 0 = parents as uint32, cast<int>() => -1 for the root (AvatarModel.cpp:36-41)
        self.parent = kintree[0].astype(np.uint32).astype(np.int32)
        assert self.parent[0] == -1
        self.baseCloud = np.ascontiguousarray(v_template.reshape(-1))             # (3V,)
        self.mesh = np.ascontiguousarray(faces)                                   # (F,3) == 3xF col-major
        self.keyClouds = np.ascontiguousarray(shapedirs.reshape(3 * V, K))        # (3V,K)
        self.weights = weights                                                    # dense copy of the J x V sparse matrix
        self.jointRegressor = jreg.T.copy()                                       # (V,J)
        # assignedJoints: (weight, joint) per vertex, weight > 1e-12, sorted descending by the pair (:74-94)
        self.assignedJoints = []
        for v in range(V):
            nz = np.nonzero(weights[v] > 1e-12)[0]
            pairs = sorted(((float(weights[v, j]), int(j)) for j in nz), reverse=True)
            self.assignedJoints.append(pairs)
        # joint shape regressor (:105-127); sparse products sum over the nonzeros in vertex order
        self.useJointShapeRegressor = True
        self.initialJointPos = np.zeros((3, J))
        self.jointShapeReg = np.zeros((3 * J, K))
        kc = self.keyClouds.reshape(V, 3, K)
        for j in range(J):
            nz = np.nonzero(jreg[j])[0]
            self.initialJointPos[:, j] = (v_template[nz] * jreg[j, nz, None]).sum(axis=0)
            self.jointShapeReg[3 * j:3 * j + 3] = np.einsum("vck,v->ck", kc[nz], jreg[j, nz])
        self.jointShapeRegBase = np.ascontiguousarray(self.initialJointPos.T.reshape(-1))
        self.posePrior = GaussianMixture()
        if pose_prior is not None:
            self.posePrior = pose_prior
        elif prior_path is not None:
            self.posePrior.load(prior_path)                                       # AvatarModel.cpp:296
        self._handle = None

    # reference accessors (include/Avatar.h:82-93)
    def numJoints(self): return int(self.parent.shape[0])
    def numPoints(self): return int(self.baseCloud.shape[0] // 3)
    def numShapeKeys(self): return int(self.keyClouds.shape[1])
    def numFaces(self): return int(self.mesh.shape[0])
    def hasMesh(self): return self.numFaces() > 0
    def hasPosePrior(self): return self.posePrior.nComps >= 0

    @property
    def handle(self):
        """avb_model* (created on first use)"""
        if self._handle is None:
            V, J, K, F = self.numPoints(), self.numJoints(), self.numShapeKeys(), self.numFaces()
            start = np.zeros(V + 1, dtype=np.int32)
            for v in range(V):
                start[v + 1] = start[v] + len(self.assignedJoints[v])
            aj = np.array([j for pairs in self.assignedJoints for (_, j) in pairs], dtype=np.int32)
            aw = np.array([w for pairs in self.assignedJoints for (w, _) in pairs], dtype=np.float64)
            keep = [self.baseCloud, self.keyClouds, self.jointShapeRegBase,
                    np.ascontiguousarray(self.jointShapeReg), np.ascontiguousarray(self.parent, dtype=np.int32),
                    np.ascontiguousarray(self.mesh, dtype=np.int32), start, aj, aw]
            d = _lib.ModelDesc()
            d.num_points, d.num_joints, d.num_shape_keys, d.num_faces = V, J, K, F
            (d.base_cloud, d.key_clouds, d.joint_shape_reg_base, d.joint_shape_reg, d.parent, d.mesh,
             d.assign_start, d.assign_joint, d.assign_weight) = [_lib.ptr(a) for a in keep]
            if self.hasPosePrior():
                g = self.posePrior
                d.gmm_components, d.gmm_dims = g.nComps, g.nDims
                d.gmm_weight, d.gmm_mean, d.gmm_cov = _lib.ptr(g.weight), _lib.ptr(g.mean), _lib.ptr(g.cov)
            h = C.c_void_p()
            _lib.check(_lib.lib.avb_model_create(C.byref(d), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        if getattr(self, "_handle", None) is not None:
            _lib.lib.avb_model_destroy(self._handle)
            self._handle = None
