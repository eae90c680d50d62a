#!/usr/bin/env python3
"""Generate the frozen synthetic fixtures under tests/golden/.

Nothing here reads /root/reference.  The reference ships no model, no pose prior and no
datasets, so the benchmark inputs are synthesised once by this script and
committed:

  tests/golden/model_synth.npz   SMPL-shaped model in the reference's npz schema
                                 (AvatarModel.cpp:26-29,35-66,99-102): v_template (6890,3) f32,
                                 shapedirs (6890,3,10) f32, f (13776,3) u32, kintree_table (2,24) u32,
                                 J_regressor (24,6890) f32, weights (6890,24) f32
  tests/golden/prior_synth.npz   8-component 69-D GMM pose prior (weights, means, covs) -- the
                                 arrays a pose_prior.txt holds (GaussianMixture.cpp:20-58) -- and the
                                 16-part joint->part map (RTree.cpp:3465-3510 .partmap semantics).

Geometry: smooth union of tapered capsules (T-pose humanoid, ~1.7 m, y up, facing +z),
meshed with marching tetrahedra, then shortest-edge collapses down to exactly 6890 vertices.
A closed genus-0 triangulation with V vertices has 2V-4 = 13776 faces, SMPL's face count.

Deterministic: numpy default_rng with fixed seeds; pure numpy/python.
"""
import heapq
import os
import sys
import numpy as np

V_TARGET = 6890
N_JOINTS = 24
N_SHAPE = 10
PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# desired rest joint positions (m); y up, +x = model's left, facing +z
JOINTS = np.array([
    [0.00, 0.00, 0.00],     # 0 pelvis
    [0.075, -0.09, 0.00],   # 1 L_hip
    [-0.075, -0.09, 0.00],  # 2 R_hip
    [0.00, 0.11, -0.02],    # 3 spine1
    [0.10, -0.47, 0.00],    # 4 L_knee
    [-0.10, -0.47, 0.00],   # 5 R_knee
    [0.00, 0.24, 0.00],     # 6 spine2
    [0.095, -0.87, -0.03],  # 7 L_ankle
    [-0.095, -0.87, -0.03], # 8 R_ankle
    [0.00, 0.31, 0.01],     # 9 spine3
    [0.10, -0.93, 0.08],    # 10 L_foot
    [-0.10, -0.93, 0.08],   # 11 R_foot
    [0.00, 0.51, -0.02],    # 12 neck
    [0.075, 0.43, 0.00],    # 13 L_collar
    [-0.075, 0.43, 0.00],   # 14 R_collar
    [0.00, 0.60, 0.02],     # 15 head
    [0.185, 0.46, -0.01],   # 16 L_shoulder
    [-0.185, 0.46, -0.01],  # 17 R_shoulder
    [0.44, 0.455, -0.02],   # 18 L_elbow
    [-0.44, 0.455, -0.02],  # 19 R_elbow
    [0.69, 0.46, -0.01],    # 20 L_wrist
    [-0.69, 0.46, -0.01],   # 21 R_wrist
    [0.775, 0.455, -0.01],  # 22 L_hand
    [-0.775, 0.455, -0.01], # 23 R_hand
], dtype=np.float64)

# joint -> body part (16 parts, every part non-empty)
PART_MAP = [0, 1, 2, 0, 3, 4, 5, 6, 7, 5, 6, 7, 8, 9, 10, 8, 9, 10, 11, 12, 13, 14, 15, 15]
# note: hands share part 15; feet share the ankle's part.
N_PARTS = 16


def capsule_sdf(p, a, b, ra, rb):
    """Signed distance (approx.) to a tapered capsule from a (radius ra) to b (radius rb)."""
    ab = b - a
    t = np.clip(((p - a) @ ab) / (ab @ ab), 0.0, 1.0)
    c = a + t[:, None] * ab
    r = ra + t * (rb - ra)
    return np.linalg.norm(p - c, axis=1) - r


def ellipsoid_sdf(p, c, r):
    q = (p - c) / r
    n = np.linalg.norm(q, axis=1)
    return (n - 1.0) * np.min(r)


def smin(a, b, k):
    h = np.clip(0.5 + 0.5 * (b - a) / k, 0.0, 1.0)
    return b + (a - b) * h - k * h * (1.0 - h)


def body_sdf(p):
    J = JOINTS
    A = lambda *v: np.array(v, dtype=np.float64)
    prims = []
    # torso: stacked ellipsoid-ish capsules
    prims.append(capsule_sdf(p, A(0, -0.04, 0.0), A(0, 0.14, -0.005), 0.125, 0.115))
    prims.append(capsule_sdf(p, A(0, 0.14, -0.005), A(0, 0.36, 0.0), 0.115, 0.125))
    prims.append(capsule_sdf(p, A(-0.09, 0.40, 0.0), A(0.09, 0.40, 0.0), 0.085, 0.085))  # shoulders yoke
    prims.append(capsule_sdf(p, A(-0.06, -0.05, 0.0), A(0.06, -0.05, 0.0), 0.105, 0.105))  # hips
    # neck + head
    prims.append(capsule_sdf(p, J[12] + A(0, -0.03, 0), J[15], 0.05, 0.05))
    prims.append(ellipsoid_sdf(p, A(0, 0.685, 0.025), A(0.085, 0.11, 0.10)))
    for s in (+1, -1):
        sx = A(s, 1, 1)
        # legs
        prims.append(capsule_sdf(p, J[1] * 1 if s > 0 else J[2], J[4] if s > 0 else J[5], 0.078, 0.055))
        prims.append(capsule_sdf(p, J[4] if s > 0 else J[5], J[7] if s > 0 else J[8], 0.055, 0.038))
        prims.append(capsule_sdf(p, (J[7] if s > 0 else J[8]) + A(0, -0.035, 0.0),
                                 A(0.10 * s, -0.945, 0.13), 0.040, 0.032))
        # arms
        prims.append(capsule_sdf(p, J[13] * sx if s < 0 else J[13], J[16] if s > 0 else J[17], 0.06, 0.05))
        prims.append(capsule_sdf(p, J[16] if s > 0 else J[17], J[18] if s > 0 else J[19], 0.05, 0.04))
        prims.append(capsule_sdf(p, J[18] if s > 0 else J[19], J[20] if s > 0 else J[21], 0.04, 0.03))
        prims.append(capsule_sdf(p, J[20] if s > 0 else J[21], A(0.84 * s, 0.455, -0.01), 0.032, 0.026))
    d = prims[0]
    for q in prims[1:]:
        d = smin(d, q, 0.03)
    return d


def marching_tets(h, lo, hi):
    nx, ny, nz = [int(np.ceil((hi[i] - lo[i]) / h)) + 1 for i in range(3)]
    xs = lo[0] + h * np.arange(nx)
    ys = lo[1] + h * np.arange(ny)
    zs = lo[2] + h * np.arange(nz)
    gx, gy, gz = np.meshgrid(xs, ys, zs, indexing="ij")
    P = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    f = body_sdf(P)
    f = np.where(np.abs(f) < 1e-7, 1e-7, f)
    gid = lambda i, j, k: (i * ny + j) * nz + k
    ii, jj, kk = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    import itertools
    tets = []
    for perm in itertools.permutations(range(3)):
        off = np.zeros((4, 3), dtype=np.int64)
        cur = np.zeros(3, dtype=np.int64)
        for s, ax in enumerate(perm):
            cur = cur.copy()
            cur[ax] = 1
            off[s + 1] = cur
        t = np.stack([gid(ii + off[s, 0], jj + off[s, 1], kk + off[s, 2]) for s in range(4)], axis=1)
        tets.append(t)
    tets = np.concatenate(tets, axis=0)
    ft = f[tets]
    inside = ft < 0
    cnt = inside.sum(axis=1)
    tris_a, tris_b = [], []  # endpoints of the grid edge each triangle corner lies on

    def emit(mask, corner_edges):
        # corner_edges: list of 3 (col_i, col_j) index arrays (per-tet local vertex indices)
        ta = np.stack([np.take_along_axis(tets[mask], ce[0][:, None], axis=1)[:, 0] for ce in corner_edges], axis=1)
        tb = np.stack([np.take_along_axis(tets[mask], ce[1][:, None], axis=1)[:, 0] for ce in corner_edges], axis=1)
        tris_a.append(ta)
        tris_b.append(tb)

    # one vertex differs from the other three
    for want_inside, c in ((True, 1), (False, 3)):
        m = cnt == c
        if not m.any():
            continue
        sel = inside[m] if want_inside else ~inside[m]
        lone = np.argmax(sel, axis=1)
        others = np.array([[j for j in range(4) if j != i] for i in range(4)])[lone]
        emit(m, [(lone, others[:, 0]), (lone, others[:, 1]), (lone, others[:, 2])])
    m = cnt == 2
    if m.any():
        ins = inside[m]
        order = np.argsort(~ins, axis=1, kind="stable")  # inside first
        a, b, c, d = order[:, 0], order[:, 1], order[:, 2], order[:, 3]
        emit(m, [(a, c), (a, d), (b, d)])
        emit(m, [(a, c), (b, d), (b, c)])
    ta = np.concatenate(tris_a, axis=0)
    tb = np.concatenate(tris_b, axis=0)
    lo_id = np.minimum(ta, tb)
    hi_id = np.maximum(ta, tb)
    key = lo_id * np.int64(P.shape[0]) + hi_id
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    faces = inv.reshape(-1, 3)
    ua = uniq // P.shape[0]
    ub = uniq % P.shape[0]
    t = f[ua] / (f[ua] - f[ub])
    verts = P[ua] + t[:, None] * (P[ub] - P[ua])
    # drop degenerate faces (repeated vertex)
    good = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    faces = faces[good]
    faces = orient_consistently(verts, faces)
    return verts, faces


def orient_consistently(verts, faces):
    """Combinatorial orientation: flood-fill across shared edges, then fix the global sign by volume."""
    faces = faces.copy()
    edge_faces = {}
    for fi, f in enumerate(faces):
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            edge_faces.setdefault((min(a, b), max(a, b)), []).append(fi)
    assert all(len(v) == 2 for v in edge_faces.values()), "marching-tets surface is not a closed manifold"
    done = np.zeros(len(faces), dtype=bool)
    ncomp = 0
    for seed in range(len(faces)):
        if done[seed]:
            continue
        ncomp += 1
        done[seed] = True
        stack = [seed]
        while stack:
            fi = stack.pop()
            f = faces[fi]
            for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
                for fj in edge_faces[(min(a, b), max(a, b))]:
                    if fj == fi or done[fj]:
                        continue
                    g = faces[fj]
                    # neighbour must traverse the shared edge in the opposite direction (b -> a)
                    dirs = ((g[0], g[1]), (g[1], g[2]), (g[2], g[0]))
                    if (a, b) in dirs:
                        faces[fj] = g[[0, 2, 1]]
                    done[fj] = True
                    stack.append(fj)
    assert ncomp == 1, "expected a single connected surface, got %d" % ncomp
    vol = np.einsum("ij,ij->i", verts[faces[:, 0]], np.cross(verts[faces[:, 1]], verts[faces[:, 2]])).sum()
    if vol < 0:
        faces = faces[:, [0, 2, 1]]
    return faces


def collapse_to(verts, faces, v_target):
    verts = [v.copy() for v in verts]
    fdict = {i: tuple(int(x) for x in f) for i, f in enumerate(faces)}
    vfaces = [set() for _ in verts]
    for fid, f in fdict.items():
        for v in f:
            vfaces[v].add(fid)
    alive = [True] * len(verts)
    nalive = len(verts)

    def elen2(a, b):
        d = verts[a] - verts[b]
        return float(d @ d)

    heap = []
    seen = set()
    for f in fdict.values():
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            k = (min(a, b), max(a, b))
            if k not in seen:
                seen.add(k)
                heapq.heappush(heap, (elen2(*k), k[0], k[1]))
    del seen

    def fnormal(f, override=None):
        p = [override[1] if (override is not None and v == override[0]) else verts[v] for v in f]
        return np.cross(p[1] - p[0], p[2] - p[0])

    while nalive > v_target and heap:
        l2, a, b = heapq.heappop(heap)
        if not (alive[a] and alive[b]):
            continue
        if abs(elen2(a, b) - l2) > 1e-15:
            continue
        Fa, Fb = vfaces[a], vfaces[b]
        shared = Fa & Fb
        if len(shared) != 2:
            continue
        Na = set(v for f in Fa for v in fdict[f]) - {a}
        Nb = set(v for f in Fb for v in fdict[f]) - {b}
        common = Na & Nb
        if len(common) != 2:
            continue
        if any(len(vfaces[c]) <= 3 for c in common):
            continue
        mid = 0.5 * (verts[a] + verts[b])
        ok = True
        for f in (Fa | Fb) - shared:
            tri = fdict[f]
            n0 = fnormal(tri)
            mv = a if a in tri else b
            n1 = fnormal(tri, (mv, mid))
            if n0 @ n1 <= 0.2 * np.sqrt((n0 @ n0) * (n1 @ n1)) or (n1 @ n1) < 1e-16:
                ok = False
                break
        if not ok:
            continue
        for f in shared:
            for v in fdict[f]:
                vfaces[v].discard(f)
            del fdict[f]
        for f in list(Fb):
            tri = fdict[f]
            fdict[f] = tuple(a if v == b else v for v in tri)
            vfaces[a].add(f)
        vfaces[b] = set()
        alive[b] = False
        verts[a] = mid
        nalive -= 1
        for nb in set(v for f in vfaces[a] for v in fdict[f]) - {a}:
            heapq.heappush(heap, (elen2(a, nb), min(a, nb), max(a, nb)))
    assert nalive == v_target, (nalive, v_target)
    remap = -np.ones(len(verts), dtype=np.int64)
    idx = [i for i, al in enumerate(alive) if al]
    remap[idx] = np.arange(len(idx))
    V = np.array([verts[i] for i in idx])
    F = np.array([[remap[v] for v in fdict[f]] for f in sorted(fdict)], dtype=np.int64)
    return V, F


def check_closed_genus0(V, F):
    E = set()
    und = {}
    for f in F:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            assert (a, b) not in E, "inconsistent orientation / non-manifold"
            E.add((a, b))
            k = (min(a, b), max(a, b))
            und[k] = und.get(k, 0) + 1
    assert all(c == 2 for c in und.values()), "not closed manifold"
    chi = len(V) - len(und) + len(F)
    assert chi == 2, chi
    vol = np.einsum("ij,ij->i", V[F[:, 0]], np.cross(V[F[:, 1]], V[F[:, 2]])).sum() / 6.0
    assert vol > 0, "normals must point outward"
    return vol


def seg_dist(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-12), 0.0, 1.0)
    c = a + t[:, None] * ab
    return np.linalg.norm(p - c, axis=1)


def main(out_dir):
    rng = np.random.default_rng(20261017)
    lo = np.array([-0.95, -1.05, -0.22])
    hi = np.array([0.95, 0.86, 0.26])
    # choose the grid step so marching tets yields somewhat more than V_TARGET vertices
    h = 0.0300
    for _ in range(12):
        verts, faces = marching_tets(h, lo, hi)
        print(f"h={h:.4f} -> V={len(verts)} F={len(faces)}", flush=True)
        if 1.25 * V_TARGET <= len(verts) <= 1.7 * V_TARGET:
            break
        h *= np.sqrt(len(verts) / (1.45 * V_TARGET))
    # keep the largest connected component only (there must be just one)
    V, F = collapse_to(verts, faces, V_TARGET)
    vol = check_closed_genus0(V, F)
    assert V.shape == (V_TARGET, 3) and F.shape == (2 * V_TARGET - 4, 3), (V.shape, F.shape)
    print("mesh ok: V", V.shape, "F", F.shape, "volume %.4f m^3" % vol)

    # ---- joint regressor: local Gaussian-weighted rings of surface vertices ----
    from scipy.optimize import nnls
    Jreg = np.zeros((N_JOINTS, V_TARGET))
    for j in range(N_JOINTS):
        d = np.linalg.norm(V - JOINTS[j], axis=1)
        nn = np.argsort(d, kind="stable")
        nn = nn[d[nn] <= d[nn[0]] + 0.07][:160]
        # non-negative weights summing to 1 whose centroid is the desired joint, regularised to be spread out
        lam, mu = 10.0, 0.05
        A = np.vstack([V[nn].T, lam * np.ones((1, len(nn))), mu * np.eye(len(nn))])
        b = np.concatenate([JOINTS[j], [lam], mu * np.full(len(nn), 1.0 / len(nn))])
        wj, _ = nnls(A, b)
        Jreg[j, nn] = wj / wj.sum()
    Jreg = Jreg.astype(np.float32)
    v_template = V.astype(np.float32)
    Jpos = Jreg.astype(np.float64) @ v_template.astype(np.float64)
    print("max |regressed joint - desired| = %.3f m" % np.abs(Jpos - JOINTS).max())

    # ---- skinning weights: Gaussian falloff of distance to bones, top-4 ----
    children = [[c for c in range(N_JOINTS) if PARENTS[c] == j] for j in range(N_JOINTS)]
    D = np.zeros((V_TARGET, N_JOINTS))
    for j in range(N_JOINTS):
        if children[j]:
            d = np.min([seg_dist(V, Jpos[j], Jpos[j] + 0.85 * (Jpos[c] - Jpos[j])) for c in children[j]], axis=0)
        else:
            ext = Jpos[j] + 0.6 * (Jpos[j] - Jpos[PARENTS[j]])
            d = seg_dist(V, Jpos[j], ext)
        D[:, j] = d
    W = np.exp(-0.5 * ((D - D.min(axis=1, keepdims=True)) / 0.035) ** 2)
    order = np.argsort(-W, axis=1, kind="stable")
    Wk = np.zeros_like(W)
    rows = np.arange(V_TARGET)[:, None]
    Wk[rows, order[:, :4]] = W[rows, order[:, :4]]
    Wk[Wk < 0.02] = 0.0
    Wk /= Wk.sum(axis=1, keepdims=True)
    weights = Wk.astype(np.float32)
    # re-normalise in f32 so rows sum to ~1
    nnz = (weights > 0).sum(axis=1)
    print("skin nnz per vertex: min %d max %d mean %.2f" % (nnz.min(), nnz.max(), nnz.mean()))
    assert nnz.max() <= 4 and nnz.min() >= 1

    # ---- shape blend shapes: smooth fields, ~1-3 cm per sigma ----
    nrm = np.zeros_like(V)
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    for c in range(3):
        np.add.at(nrm, F[:, c], fn)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    shapedirs = np.zeros((V_TARGET, 3, N_SHAPE))
    yc = V[:, 1] + 0.1
    shapedirs[:, :, 0] = 0.035 * V * np.array([0.6, 1.0, 0.6])          # overall size / height
    shapedirs[:, :, 1] = 0.012 * nrm                                      # girth
    shapedirs[:, :, 2] = 0.010 * nrm * np.tanh(4.0 * yc)[:, None]         # upper vs lower body mass
    shapedirs[:, 1, 3] = 0.020 * np.maximum(-V[:, 1], 0.0)                # leg length
    shapedirs[:, 0, 4] = 0.030 * np.sign(V[:, 0]) * np.maximum(np.abs(V[:, 0]) - 0.15, 0.0)  # arm length
    for m in range(5, N_SHAPE):
        fr = rng.uniform(2.0, 7.0, size=3)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        amp = 0.008 * (0.8 ** (m - 5))
        field = np.sin(V @ np.diag(fr) + ph).prod(axis=1)
        shapedirs[:, :, m] = amp * nrm * field[:, None]
    shapedirs = shapedirs.astype(np.float32)

    kintree = np.zeros((2, N_JOINTS), dtype=np.uint32)
    kintree[0] = np.array(PARENTS, dtype=np.int64).astype(np.uint32)  # -1 -> 4294967295
    kintree[1] = np.arange(N_JOINTS)

    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "model_synth.npz"),
                        v_template=v_template, shapedirs=shapedirs, f=F.astype(np.uint32),
                        kintree_table=kintree, J_regressor=Jreg, weights=weights)

    # ---- GMM pose prior: C=8, D=69 ----
    C, Dm = 8, 3 * (N_JOINTS - 1)
    gw = rng.uniform(0.5, 1.5, size=C)
    gw /= gw.sum()
    means = rng.normal(0.0, 0.12, size=(C, Dm))
    covs = np.zeros((C, Dm, Dm))
    for c in range(C):
        A = rng.normal(0.0, 1.0, size=(Dm, 12)) * 0.08
        sig = rng.uniform(0.2, 0.5, size=Dm)
        covs[c] = A @ A.T + np.diag(sig ** 2)
        covs[c] = 0.5 * (covs[c] + covs[c].T)
    np.savez_compressed(os.path.join(out_dir, "prior_synth.npz"), weights=gw, means=means, covs=covs,
                        part_map=np.array(PART_MAP, dtype=np.int32), num_parts=np.int32(N_PARTS))
    for fn_ in ("model_synth.npz", "prior_synth.npz"):
        print(fn_, os.path.getsize(os.path.join(out_dir, fn_)) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
