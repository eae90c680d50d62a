// avb_kernels.cu -- hand-written sm_100a kernels of the SMPL-to-point-cloud fit.
//
//   pose_visibility_kernel : Avatar::update (Avatar.cpp:22-75) + back-face visibility
//                            (AvatarOptimizer.cpp:1349-1367) + per-part compaction of the visible
//                            model vertices (AvatarOptimizer.cpp:860-880).  One CTA per frame.
//   nn_kernel              : findNN(..., invert=true) (AvatarOptimizer.cpp:841-920): exact fp64
//                            1-NN of every data point among the visible model vertices of its body
//                            part.  The frame's compacted model cloud is staged into shared memory
//                            with one TMA bulk copy (cp.async.bulk + mbarrier); data points stream
//                            from HBM.  Also reduces the correspondences per model vertex.
//   lm_fit_kernel          : the whole inner solve of one ICP iteration for one frame in one
//                            persistent CTA: residuals + analytic Jacobians (AvatarOptimizer.cpp:505-582),
//                            priors (:661-692, :708-723), J^T J / J^T r, damped Cholesky solve,
//                            retraction (:123-143), step control.  No host round trips.
//
// Numerics: geometry, NN distances, residuals, gradient, cost and the linear solve are fp64;
// the Gauss-Newton matrix is accumulated from fp32 Jacobian rows (it only steers the step).
#include "avb_device.cuh"
#include "avb_kernels.h"

#include <math.h>

namespace avb {

// ---------------------------------------------------------------------------------------------
// joint tables (per frame, in shared memory)
// ---------------------------------------------------------------------------------------------
struct Tables {
    double* Jr;   // [J][3]   shaped rest joint positions (jointShapeRegBase + jointShapeReg w)
    double* Rq;   // [J][9]   local rotations R(q_j)
    double* G;    // [J][9]   global rotations  R(-1, j)            (AvatarOptimizer.cpp:303-315)
    double* pos;  // [J][3]   global joint positions t(-1, j)
    double* tau;  // [J][3]   skinning translation pos_j - G_j Jr_j  (Avatar.cpp:59-64)
    double* Hj;   // [J][3][K] accumulated shape deltas H[j]          (AvatarOptimizer.cpp:318-324)
    double* C;    // [J][3][K] H[j] - G_j S_j
};
__host__ __device__ inline int tables_doubles(int J, int K, bool with_shape) {
    return J * (3 + 9 + 9 + 3 + 3) + (with_shape ? 2 * J * 3 * K : 0);
}
__device__ inline Tables carve_tables(double* base, int J, int K, bool with_shape) {
    Tables T;
    T.Jr = base; base += 3 * J;
    T.Rq = base; base += 9 * J;
    T.G = base; base += 9 * J;
    T.pos = base; base += 3 * J;
    T.tau = base; base += 3 * J;
    T.Hj = with_shape ? base : nullptr; base += with_shape ? 3 * J * K : 0;
    T.C = with_shape ? base : nullptr;
    return T;
}

// CTA-wide.  xs = [p | q | w] in shared memory.
__device__ void build_tables(const DevModel& M, const double* xs, Tables T, bool with_shape) {
    const int J = M.J, K = M.K, tid = threadIdx.x, nt = blockDim.x;
    const double* w = xs + 3 + 4 * J;
    for (int i = tid; i < 3 * J; i += nt) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += M.jreg[i * K + k] * w[k];
        T.Jr[i] = M.jbase[i] + s;
    }
    for (int j = tid; j < J; j += nt) quat_to_rot(xs + 3 + 4 * j, T.Rq + 9 * j);
    __syncthreads();
    for (int d = 0; d <= M.max_depth; ++d) {
        for (int j = tid; j < J; j += nt) {
            if (M.depth[j] != d) continue;
            const int pa = M.parent[j];
            double* Gj = T.G + 9 * j;
            const double* R = T.Rq + 9 * j;
            if (pa < 0) {
                for (int e = 0; e < 9; ++e) Gj[e] = R[e];
                for (int c = 0; c < 3; ++c) T.pos[3 * j + c] = xs[c];
            } else {
                const double* Gp = T.G + 9 * pa;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        Gj[3 * r + c] = Gp[3 * r] * R[c] + Gp[3 * r + 1] * R[3 + c] + Gp[3 * r + 2] * R[6 + c];
                const double v0 = T.Jr[3 * j] - T.Jr[3 * pa], v1 = T.Jr[3 * j + 1] - T.Jr[3 * pa + 1],
                             v2 = T.Jr[3 * j + 2] - T.Jr[3 * pa + 2];
                for (int r = 0; r < 3; ++r)
                    T.pos[3 * j + r] = Gp[3 * r] * v0 + Gp[3 * r + 1] * v1 + Gp[3 * r + 2] * v2 + T.pos[3 * pa + r];
            }
        }
        __syncthreads();
    }
    for (int j = tid; j < J; j += nt) {
        const double* Gj = T.G + 9 * j;
        for (int r = 0; r < 3; ++r)
            T.tau[3 * j + r] = T.pos[3 * j + r] -
                               (Gj[3 * r] * T.Jr[3 * j] + Gj[3 * r + 1] * T.Jr[3 * j + 1] + Gj[3 * r + 2] * T.Jr[3 * j + 2]);
    }
    if (with_shape) {
        const int per = 3 * K;
        for (int i = tid; i < J * per; i += nt)
            if (M.depth[i / per] == 0) T.Hj[i] = 0.0;
        __syncthreads();
        for (int d = 1; d <= M.max_depth; ++d) {
            for (int i = tid; i < J * per; i += nt) {
                const int j = i / per;
                if (M.depth[j] != d) continue;
                const int r = (i % per) / K, m = i % K, pa = M.parent[j];
                const double* Gp = T.G + 9 * pa;
                const double* sp = M.Sp + (size_t)j * per;
                T.Hj[i] = Gp[3 * r] * sp[m] + Gp[3 * r + 1] * sp[K + m] + Gp[3 * r + 2] * sp[2 * K + m] +
                          T.Hj[pa * per + r * K + m];
            }
            __syncthreads();
        }
        for (int i = tid; i < J * per; i += nt) {
            const int j = i / per, r = (i % per) / K, m = i % K;
            const double* Gj = T.G + 9 * j;
            const double* S = M.jreg + (size_t)3 * j * K;
            T.C[i] = T.Hj[i] - (Gj[3 * r] * S[m] + Gj[3 * r + 1] * S[K + m] + Gj[3 * r + 2] * S[2 * K + m]);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// K1: forward SMPL + visibility + per-part compaction.  grid = batch, block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pose_visibility_kernel(DevModel M, DevParts Pt, PoseArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int V = M.V, J = M.J, K = M.K;
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* tb = xs + ((M.nx + 1) & ~1);
    Tables T = carve_tables(tb, J, K, false);
    int* iscr = reinterpret_cast<int*>(tb + tables_doubles(J, K, false));  // [64]
    uint8_t* vis = reinterpret_cast<uint8_t*>(iscr + 64);                    // [V]

    for (int i = tid; i < M.nx; i += nt) xs[i] = a.x[(size_t)f * M.nx + i];
    __syncthreads();
    build_tables(M, xs, T, false);

    double* cloud = a.cloud + (size_t)f * 3 * V;
    const double* w = xs + 3 + 4 * J;
    for (int v = tid; v < V; v += nt) {
        // shape blend (Avatar.cpp:26) then LBS with the assignedJoints weights (AvatarOptimizer.cpp:507-514)
        const float* sd = M.sd + (size_t)v * 3 * K;
        double v0[3];
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)sd[c * K + k] * w[k];
            v0[c] = M.vt[3 * (size_t)v + c] + s;
        }
        double x0 = 0, x1 = 0, x2 = 0;
        const int n = M.sk_n[v];
        for (int q = 0; q < n; ++q) {
            const int k = M.sk_j[4 * (size_t)v + q];
            const double wt = M.sk_w[4 * (size_t)v + q];
            const double* G = T.G + 9 * k;
            x0 += wt * (G[0] * v0[0] + G[1] * v0[1] + G[2] * v0[2] + T.tau[3 * k]);
            x1 += wt * (G[3] * v0[0] + G[4] * v0[1] + G[5] * v0[2] + T.tau[3 * k + 1]);
            x2 += wt * (G[6] * v0[0] + G[7] * v0[1] + G[8] * v0[2] + T.tau[3 * k + 2]);
        }
        cloud[3 * (size_t)v] = x0;
        cloud[3 * (size_t)v + 1] = x1;
        cloud[3 * (size_t)v + 2] = x2;
    }
    if (a.joint_pos)
        for (int i = tid; i < 3 * J; i += nt) a.joint_pos[(size_t)f * 3 * J + i] = T.pos[i];
    if (a.joint_trans)  // 3x4 column-major per joint: [G_j | tau_j] (Avatar.cpp:59-64)
        for (int i = tid; i < 12 * J; i += nt) {
            const int j = i / 12, e = i % 12, c = e / 3, r = e % 3;
            a.joint_trans[(size_t)f * 12 * J + i] = (c < 3) ? T.G[9 * j + 3 * r + c] : T.tau[3 * j + r];
        }
    if (!a.do_visibility) return;

    // ---- back-face visibility (AvatarOptimizer.cpp:1349-1367) ----
    for (int v = tid; v < V; v += nt) vis[v] = a.enable_occlusion ? 0 : 1;
    __threadfence_block();
    __syncthreads();
    if (a.enable_occlusion) {
        for (int t = tid; t < M.F; t += nt) {
            const int i1 = M.faces[3 * t], i2 = M.faces[3 * t + 1], i3 = M.faces[3 * t + 2];
            const double p1x = cloud[3 * (size_t)i1], p1y = cloud[3 * (size_t)i1 + 1];
            const double p2x = cloud[3 * (size_t)i2], p2y = cloud[3 * (size_t)i2 + 1];
            const double p3x = cloud[3 * (size_t)i3], p3y = cloud[3 * (size_t)i3 + 1];
            // ((p2 - p1).cross(p1 - p3)).z(); separate products, no fma contraction, as the CPU compiles it
            const double ax = p2x - p1x, ay = p2y - p1y, bx = p1x - p3x, by = p1y - p3y;
            const double z = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
            if (z > 1e-4) vis[i1] = vis[i2] = vis[i3] = 1;
        }
        __syncthreads();
    }
    uint8_t* gvis = a.visible + (size_t)f * V;
    for (int v = tid; v < V; v += nt) gvis[v] = vis[v];

    // ---- compaction per part, ascending vertex id inside a part (AvatarOptimizer.cpp:860-880) ----
    int* pv_idx = a.pv_idx + (size_t)f * V;
    double* pv_xyz = a.pv_xyz + (size_t)f * a.pv_stride;
    int* pv_start = a.pv_start + (size_t)f * (Pt.numParts + 1);
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int base = 0;
    for (int i0 = 0; i0 < V; i0 += nt) {
        const int i = i0 + tid;
        int v = -1, flag = 0;
        if (i < V) {
            v = Pt.part_verts[i];
            flag = vis[v];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int wpre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) iscr[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int q = 0; q < nw; ++q) {
            if (q < wid) woff += iscr[q];
            tot += iscr[q];
        }
        const int pos = base + woff + wpre;  // exclusive prefix of visible flags before element i
        if (i < V) {
            int p = Pt.first_part_at[i];
            if (p >= 0)
                for (; p < Pt.numParts && Pt.part_start[p] == i; ++p) pv_start[p] = pos;
            if (flag) {
                pv_idx[pos] = v;
                pv_xyz[3 * (size_t)pos] = cloud[3 * (size_t)v];
                pv_xyz[3 * (size_t)pos + 1] = cloud[3 * (size_t)v + 1];
                pv_xyz[3 * (size_t)pos + 2] = cloud[3 * (size_t)v + 2];
            }
        }
        base += tot;
        __syncthreads();
    }
    if (tid == 0) {
        int p = Pt.first_part_at[V];
        if (p >= 0)
            for (; p < Pt.numParts && Pt.part_start[p] == V; ++p) pv_start[p] = base;
        pv_start[Pt.numParts] = base;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: exact fp64 1-NN, data -> visible model vertices of the same part.  block = 512
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1)
nn_kernel(DevParts Pt, NNArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int s_start[kMaxParts + 1];
    __shared__ double s_wsum[16];
    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    const int f = a.chunk_frame[c];
    const long long begin = a.chunk_begin[c];
    const int count = a.chunk_count[c];
    double* mxyz = reinterpret_cast<double*>(smem_raw);

    const int* pv_start = a.pv_start + (size_t)f * (Pt.numParts + 1);
    if (tid <= Pt.numParts) s_start[tid] = pv_start[tid];
    const int nvis = pv_start[Pt.numParts];
    // one TMA bulk copy of the frame's compacted model cloud (nvis * 24 B, rounded up to 16 B)
    const uint32_t bytes = (uint32_t)(((size_t)nvis * 24 + 15) & ~(size_t)15);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && bytes > 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes)
                     : "memory");
        const double* src = a.pv_xyz + (size_t)f * a.pv_stride;
        uint32_t off = 0;
        while (off < bytes) {  // <= 64 KiB per bulk copy
            const uint32_t n = min(bytes - off, 65536u);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(reinterpret_cast<unsigned char*>(mxyz) + off)),
                "l"(reinterpret_cast<const unsigned char*>(src) + off), "r"(n), "r"(smem_u32(&mbar))
                : "memory");
            off += n;
        }
    }
    if (bytes > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(&mbar))
                : "memory");
        }
    }
    __syncthreads();

    const int* pv_idx = a.pv_idx + (size_t)f * a.V;
    int* cnt = a.cnt + (size_t)f * a.V;
    unsigned long long* sum = a.sum + (size_t)f * 3 * a.V;
    const int lane = tid & 31, wid = tid >> 5;
    // passes of 512 points = two deterministic |d|^2 blocks of kQBlock (256) points
    for (int base = 0; base < count; base += 512) {
        const int li = base + tid;
        double q2 = 0.0;
        if (li < count) {
            const long long gi = begin + li;
            const double q0 = a.data[3 * gi], q1 = a.data[3 * gi + 1], qz = a.data[3 * gi + 2];
            const int label = a.labels[gi];
            const bool lab_ok = label >= 0 && label < Pt.numParts;  // out of range is UB in the reference (:1279)
            if (!lab_ok) atomicOr(a.range_flag + f, 2);
            const int s = lab_ok ? s_start[label] : 0, e = lab_ok ? s_start[label + 1] : 0;
            double best = 1.79769313486231570e308;
            int bi = -1;
#pragma unroll 4
            for (int k = s; k < e; ++k) {
                // nanoflann L2_Simple: sum_c (a_c - b_c)^2, c = 0,1,2 (nanoflann.hpp:423-445); strict '<'
                const double d0 = q0 - mxyz[3 * k], d1 = q1 - mxyz[3 * k + 1], d2 = qz - mxyz[3 * k + 2];
                const double dist = __fma_rn(d2, d2, __fma_rn(d1, d1, __dmul_rn(d0, d0)));
                if (dist < best) {
                    best = dist;
                    bi = k;
                }
            }
            int vtx = -1;
            if (bi >= 0) {
                vtx = pv_idx[bi];
                const bool ok = fabs(q0) < kCoordLimit && fabs(q1) < kCoordLimit && fabs(qz) < kCoordLimit;
                if (ok) {
                    atomicAdd(&cnt[vtx], 1);
                    atomicAdd(&sum[3 * (size_t)vtx], (unsigned long long)__double2ll_rn(q0 * kFixScale));
                    atomicAdd(&sum[3 * (size_t)vtx + 1], (unsigned long long)__double2ll_rn(q1 * kFixScale));
                    atomicAdd(&sum[3 * (size_t)vtx + 2], (unsigned long long)__double2ll_rn(qz * kFixScale));
                    // |d|^2 of the fixed-point-rounded point, so that cost and gradient see the same data
                    const double r0 = (double)__double2ll_rn(q0 * kFixScale) * kFixInv,
                                 r1 = (double)__double2ll_rn(q1 * kFixScale) * kFixInv,
                                 r2 = (double)__double2ll_rn(qz * kFixScale) * kFixInv;
                    q2 = r0 * r0 + r1 * r1 + r2 * r2;
                } else {
                    vtx = -1;
                    atomicOr(a.range_flag + f, 1);
                }
            }
            a.nn_idx[gi] = vtx;
        }
        // deterministic per-256-point partial of sum |d|^2
        q2 = warp_sum(q2);
        __syncthreads();
        if (lane == 0) s_wsum[wid] = q2;
        __syncthreads();
        if (tid < 2) {
            const int blk = (base / kQBlock) + tid;  // block index inside this chunk
            if (blk * kQBlock < count) {
                double s = 0;
                for (int q = 0; q < 8; ++q) s += s_wsum[8 * tid + q];
                a.qpart[a.chunk_qblock[c] + blk] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4-K7: persistent per-frame Levenberg-Marquardt kernel.  grid = batch, block = 256
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 64;   // matched vertices per Jacobian tile
constexpr int kLmThreads = 256;

struct LmSmem {
    double* xs;      // [nx]  current point
    double* xt;      // [nx]  trial point
    double* tb;      // joint tables
    double* Hs;      // [P][P] Gauss-Newton matrix being accumulated / transformed
    double* gs;      // [P]   gradient being accumulated
    double* gcur;    // [P]   gradient at the current point
    double* delta;   // [P]
    double* vec;     // [P]   scratch vector
    double* aa;      // [D]   axis-angle pose vector for the prior
    double* ycomp;   // [C][D] Sigma_c^-1 (x - mu_c)
    double* scr;     // [64]  reduction scratch / scalars
    double* rho;     // [3*kTile] residual sums / sqrt(count)
    float* A;        // [3*kTile][lda]  Jacobian tile (rows scaled by sqrt(count)); aliased by the Cholesky workspace
    unsigned short* mlist;  // [V] matched vertices grouped by column group
    int* gcount;     // [kMaxGroups+1] starts into mlist
    int* iscr;       // [64]
};

__host__ __device__ inline size_t lm_smem_bytes(int V, int J, int K, int C) {
    const int P = 3 + 3 * J + K, nx = 3 + 4 * J + K, D = 3 * (J - 1);
    const int lda = ((P + 7) & ~7) + 4;
    size_t d = 0;
    d += 2 * ((nx + 1) & ~1);
    d += tables_doubles(J, K, true);
    d += (size_t)P * P + 4 * ((P + 1) & ~1) + ((D + 1) & ~1) + (size_t)(C > 0 ? C : 1) * ((D + 1) & ~1) + 64 + 3 * kTile;
    size_t bytes = d * 8;
    size_t a_bytes = (size_t)3 * kTile * lda * 4;
    size_t w_bytes = (size_t)P * P * 8;  // Cholesky workspace aliases the tile
    bytes += (a_bytes > w_bytes ? a_bytes : w_bytes);
    bytes += ((size_t)V * 2 + 15) & ~(size_t)15;
    bytes += (kMaxGroups + 1 + 64) * 4;
    return bytes + 64;
}

__device__ inline LmSmem carve_lm(unsigned char* raw, const DevModel& M) {
    LmSmem S;
    const int P = M.P, nx = M.nx, D = 3 * (M.J - 1), C = M.gmmC > 0 ? M.gmmC : 1;
    const int lda = ((P + 7) & ~7) + 4;
    double* d = reinterpret_cast<double*>(raw);
    S.xs = d; d += (nx + 1) & ~1;
    S.xt = d; d += (nx + 1) & ~1;
    S.tb = d; d += tables_doubles(M.J, M.K, true);
    S.Hs = d; d += (size_t)P * P;
    S.gs = d; d += (P + 1) & ~1;
    S.gcur = d; d += (P + 1) & ~1;
    S.delta = d; d += (P + 1) & ~1;
    S.vec = d; d += (P + 1) & ~1;
    S.aa = d; d += (D + 1) & ~1;
    S.ycomp = d; d += (size_t)C * ((D + 1) & ~1);
    S.scr = d; d += 64;
    S.rho = d; d += 3 * kTile;
    unsigned char* b = reinterpret_cast<unsigned char*>(d);
    b = reinterpret_cast<unsigned char*>(((uintptr_t)b + 15) & ~(uintptr_t)15);
    S.A = reinterpret_cast<float*>(b);
    size_t a_bytes = (size_t)3 * kTile * lda * 4, w_bytes = (size_t)P * P * 8;
    b += (a_bytes > w_bytes ? a_bytes : w_bytes);
    S.mlist = reinterpret_cast<unsigned short*>(b);
    b += ((size_t)M.V * 2 + 15) & ~(size_t)15;
    S.gcount = reinterpret_cast<int*>(b);
    S.iscr = S.gcount + kMaxGroups + 1;
    return S;
}

// One Jacobian row-triple per thread: position, residual statistics and the tangent Jacobian in
// "global-frame rotation" coordinates eta_j = G_parent(j) delta_j (AvatarOptimizer.cpp:505-582 in
// closed form: block_j = R(-1,parent j) dRot Lq_j = -2 [y_j]x G_parent(j)).
__device__ __forceinline__ double vertex_rows(const DevModel& M, const Tables& T, const double* w, int v, int cntv,
                                              const unsigned long long* sumv, const int* gj, int nj, int lda,
                                              float* Arow /* 3 rows */, double* rho3) {
    const int K = M.K;
    const float* sd = M.sd + (size_t)v * 3 * K;
    double v0[3];
    for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)sd[c * K + k] * w[k];
        v0[c] = M.vt[3 * (size_t)v + c] + s;
    }
    const int n = M.sk_n[v];
    double xk[AVB_MAX_ASSIGN_][3], wk[AVB_MAX_ASSIGN_];
    int jk[AVB_MAX_ASSIGN_];
    uint32_t mk[AVB_MAX_ASSIGN_];
    double x[3] = {0, 0, 0};
    double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
        if (q < n) {
            const int k = M.sk_j[4 * (size_t)v + q];
            const double wt = M.sk_w[4 * (size_t)v + q];
            const double* G = T.G + 9 * k;
            jk[q] = k;
            wk[q] = wt;
            mk[q] = M.anc_mask[k];
            for (int r = 0; r < 3; ++r) {
                xk[q][r] = G[3 * r] * v0[0] + G[3 * r + 1] * v0[1] + G[3 * r + 2] * v0[2] + T.tau[3 * k + r];
                x[r] += wt * xk[q][r];
            }
            for (int e = 0; e < 9; ++e) B[e] += wt * G[e];
        } else {
            jk[q] = 0; wk[q] = 0; mk[q] = 0;
            xk[q][0] = xk[q][1] = xk[q][2] = 0;
        }
    }
    const double cn = (double)cntv;
    double s3[3], r3[3];
    for (int c = 0; c < 3; ++c) {
        s3[c] = (double)(long long)sumv[c] * kFixInv;
        r3[c] = cn * x[c] - s3[c];
    }
    // sum_i |x - d_i|^2 - sum_i |d_i|^2 = x . (c x - 2 s)
    const double costv = x[0] * (r3[0] - s3[0]) + x[1] * (r3[1] - s3[1]) + x[2] * (r3[2] - s3[2]);
    const double sc = sqrt(cn), isc = 1.0 / sc;
    for (int c = 0; c < 3; ++c) rho3[c] = r3[c] * isc;
    const float scf = (float)sc;
    float* A0 = Arow;
    float* A1 = Arow + lda;
    float* A2 = Arow + 2 * lda;
    // root translation: identity (AvatarOptimizer.cpp:477-481)
    A0[0] = scf; A0[1] = 0; A0[2] = 0;
    A1[0] = 0; A1[1] = scf; A1[2] = 0;
    A2[0] = 0; A2[1] = 0; A2[2] = scf;
    for (int gi = 0; gi < nj; ++gi) {
        const int j = gj[gi];
        double y0 = 0, y1 = 0, y2 = 0, W = 0;
#pragma unroll
        for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
            if ((mk[q] >> j) & 1u) {
                W += wk[q];
                y0 += wk[q] * xk[q][0];
                y1 += wk[q] * xk[q][1];
                y2 += wk[q] * xk[q][2];
            }
        }
        y0 = (y0 - W * T.pos[3 * j]) * (2.0 * sc);
        y1 = (y1 - W * T.pos[3 * j + 1]) * (2.0 * sc);
        y2 = (y2 - W * T.pos[3 * j + 2]) * (2.0 * sc);
        const float f0 = (float)y0, f1 = (float)y1, f2 = (float)y2;
        const int c0 = 3 + 3 * gi;
        // -2 [y]x
        A0[c0] = 0;    A0[c0 + 1] = f2;  A0[c0 + 2] = -f1;
        A1[c0] = -f2;  A1[c0 + 1] = 0;   A1[c0 + 2] = f0;
        A2[c0] = f1;   A2[c0 + 1] = -f0; A2[c0 + 2] = 0;
    }
    // shape: sum_k w_k (G_k (Delta_v - S_k) + H_k) = B Delta_v + sum_k w_k C_k  (AvatarOptimizer.cpp:568-580)
    const int cs = 3 + 3 * nj;
    for (int m = 0; m < K; ++m) {
        const double d0 = sd[m], d1 = sd[K + m], d2 = sd[2 * K + m];
        double e0 = B[0] * d0 + B[1] * d1 + B[2] * d2;
        double e1 = B[3] * d0 + B[4] * d1 + B[5] * d2;
        double e2 = B[6] * d0 + B[7] * d1 + B[8] * d2;
#pragma unroll
        for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
            if (q < n) {
                const double* Cq = T.C + (size_t)jk[q] * 3 * K;
                e0 += wk[q] * Cq[m];
                e1 += wk[q] * Cq[K + m];
                e2 += wk[q] * Cq[2 * K + m];
            }
        }
        A0[cs + m] = (float)(e0 * sc);
        A1[cs + m] = (float)(e1 * sc);
        A2[cs + m] = (float)(e2 * sc);
    }
    return costv;
}

// Objective at xs: fills S.Hs (P x P, reference tangent coordinates, priors included), S.gs, returns cost.
// AccT = double: J^T J accumulated in fp64 from the fp32 Jacobian rows (default, parity path);
// AccT = float : fp32 accumulation (AVB_JTJ_FP32, faster, ~1e-4 parameter drift over 10 iterations).
template <typename AccT>
__device__ double lm_evaluate(const DevModel& M, const DevParts& Pt, const LmArgs& a, LmSmem& S, const double* xs,
                              int f, double Qsum, double sbp, double sbs) {
    const int tid = threadIdx.x, P = M.P, J = M.J, K = M.K;
    const int lda = ((P + 7) & ~7) + 4;
    Tables T = carve_tables(S.tb, J, K, true);
    build_tables(M, xs, T, true);
    for (int i = tid; i < P * P; i += kLmThreads) S.Hs[i] = 0.0;
    for (int i = tid; i < P; i += kLmThreads) S.gs[i] = 0.0;
    __syncthreads();
    const double* w = xs + 3 + 4 * J;
    const int* cnt = a.cnt + (size_t)f * M.V;
    const unsigned long long* sum = a.sum + (size_t)f * 3 * M.V;
    double cost_acc = 0.0;

    for (int g = 0; g < Pt.numGroups; ++g) {
        const int nj = Pt.gnj[g];
        const int* gj = Pt.gjoints + g * kMaxJ;
        const int L = 3 + 3 * nj + K;
        const int Lp = (L + 7) & ~7;
        const int ntile = Lp >> 3, ntri = ntile * (ntile + 1) / 2;
        const int nrg = kLmThreads / ntri > 0 ? kLmThreads / ntri : 1;
        // syrk role: (tile pair, row group)
        int ti = -1, tj = -1, rg = 0;
        if (tid < ntri * nrg) {
            rg = tid / ntri;
            int t = tid % ntri;
            ti = 0;
            int rowlen = ntile;
            while (t >= rowlen) {
                t -= rowlen;
                ++ti;
                --rowlen;
            }
            tj = ti + t;
        }
        // gradient role: (column, row chunk)
        const int nch = kLmThreads / Lp;
        const int gcol = tid % Lp, gch = tid / Lp;
        AccT acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = AccT(0);
        double gacc = 0.0;

        const int mbeg = S.gcount[g], mend = S.gcount[g + 1];
        for (int t0 = mbeg; t0 < mend; t0 += kTile) {
            const int nv = min(kTile, mend - t0);
            // step 1: Jacobian rows of the tile (zero the pad columns once per row)
            if (tid < nv) {
                const int v = S.mlist[t0 + tid];
                float* Arow = S.A + (size_t)3 * tid * lda;
                for (int c = L; c < Lp; ++c) Arow[c] = Arow[lda + c] = Arow[2 * lda + c] = 0.f;
                cost_acc += vertex_rows(M, T, w, v, cnt[v], sum + 3 * (size_t)v, gj, nj, lda, Arow, S.rho + 3 * tid);
            }
            __syncthreads();
            const int nrows = 3 * nv;
            // step 2a: upper-triangular 8x8 register tiles of A^T A
            if (ti >= 0) {
                const float* Ab = S.A + 8 * ti;
                const float* Bb = S.A + 8 * tj;
                for (int r = rg; r < nrows; r += nrg) {
                    const float4 a0 = *reinterpret_cast<const float4*>(Ab + (size_t)r * lda);
                    const float4 a1 = *reinterpret_cast<const float4*>(Ab + (size_t)r * lda + 4);
                    const float4 b0 = *reinterpret_cast<const float4*>(Bb + (size_t)r * lda);
                    const float4 b1 = *reinterpret_cast<const float4*>(Bb + (size_t)r * lda + 4);
                    const AccT av[8] = {AccT(a0.x), AccT(a0.y), AccT(a0.z), AccT(a0.w),
                                        AccT(a1.x), AccT(a1.y), AccT(a1.z), AccT(a1.w)};
                    const AccT bv[8] = {AccT(b0.x), AccT(b0.y), AccT(b0.z), AccT(b0.w),
                                        AccT(b1.x), AccT(b1.y), AccT(b1.z), AccT(b1.w)};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                }
            }
            // step 2b: gradient J^T r in fp64
            if (gch < nch) {
                for (int r = gch; r < nrows; r += nch) gacc += (double)S.A[(size_t)r * lda + gcol] * S.rho[r];
            }
            __syncthreads();
        }
        // flush the group's partial sums into the global-column matrix
        auto colmap = [&](int c) -> int {
            if (c < 3) return c;
            if (c < 3 + 3 * nj) return 3 + 3 * gj[(c - 3) / 3] + (c - 3) % 3;
            if (c < L) return 3 + 3 * J + (c - 3 - 3 * nj);
            return -1;
        };
        if (ti >= 0 && mend > mbeg) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ci = colmap(8 * ti + i);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cj = colmap(8 * tj + j);
                    if (ci >= 0 && cj >= 0 && (ti != tj || i <= j)) atomicAdd(&S.Hs[(size_t)ci * P + cj], (double)acc[i][j]);
                }
            }
        }
        if (gch < nch && mend > mbeg) {
            const int cg = colmap(gcol);
            if (cg >= 0) atomicAdd(&S.gs[cg], gacc);
        }
        __syncthreads();
    }
    double cost = 0.5 * (block_sum(cost_acc, S.scr) + Qsum);
    // mirror to the lower triangle
    for (int i = tid; i < P * P; i += kLmThreads) {
        const int r = i / P, c = i % P;
        if (r > c) S.Hs[i] = S.Hs[(size_t)c * P + r];
    }
    __syncthreads();
    // eta -> delta coordinates: H = T^T Ht T, g = T^T gt with T_j = G_parent(j) (identity for the root)
    for (int i = tid; i < P * (J - 1); i += kLmThreads) {  // rows: block j, column c
        const int j = 1 + i / P, c = i % P;
        const double* Gp = T.G + 9 * M.parent[j];
        double* h = S.Hs + (size_t)(3 + 3 * j) * P + c;
        const double h0 = h[0], h1 = h[P], h2 = h[2 * P];
        h[0] = Gp[0] * h0 + Gp[3] * h1 + Gp[6] * h2;
        h[P] = Gp[1] * h0 + Gp[4] * h1 + Gp[7] * h2;
        h[2 * P] = Gp[2] * h0 + Gp[5] * h1 + Gp[8] * h2;
    }
    __syncthreads();
    for (int i = tid; i < P * (J - 1); i += kLmThreads) {  // columns: block j, row r
        const int j = 1 + i / P, r = i % P;
        const double* Gp = T.G + 9 * M.parent[j];
        double* h = S.Hs + (size_t)r * P + 3 + 3 * j;
        const double h0 = h[0], h1 = h[1], h2 = h[2];
        h[0] = h0 * Gp[0] + h1 * Gp[3] + h2 * Gp[6];
        h[1] = h0 * Gp[1] + h1 * Gp[4] + h2 * Gp[7];
        h[2] = h0 * Gp[2] + h1 * Gp[5] + h2 * Gp[8];
    }
    for (int j = 1 + tid; j < J; j += kLmThreads) {
        const double* Gp = T.G + 9 * M.parent[j];
        double* gg = S.gs + 3 + 3 * j;
        const double g0 = gg[0], g1 = gg[1], g2 = gg[2];
        gg[0] = Gp[0] * g0 + Gp[3] * g1 + Gp[6] * g2;
        gg[1] = Gp[1] * g0 + Gp[4] * g1 + Gp[7] * g2;
        gg[2] = Gp[2] * g0 + Gp[5] * g1 + Gp[8] * g2;
    }
    __syncthreads();
    // ---- pose prior (AvatarOptimizer.cpp:661-692, GaussianMixture.cpp:95-114) ----
    if (sbp > 0.0 && M.gmmC > 0) {
        const int D = M.gmmD, C = M.gmmC, Dp = (D + 1) & ~1;
        for (int j = 1 + tid; j < J; j += kLmThreads) {  // Eigen AngleAxisd(Quaterniond)
            const double* q = xs + 3 + 4 * j;
            double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            double s = 0.0;
            if (n != 0.0) {
                const double ang = 2.0 * atan2(n, fabs(q[3]));
                if (q[3] < 0) n = -n;
                s = ang / n;
            }
            S.aa[3 * (j - 1)] = q[0] * s;
            S.aa[3 * (j - 1) + 1] = q[1] * s;
            S.aa[3 * (j - 1) + 2] = q[2] * s;
        }
        __syncthreads();
        for (int i = tid; i < C * D; i += kLmThreads) {  // y_c = Sigma_c^-1 (x - mu_c)
            const int c = i / D, r = i % D;
            const double* Pm = M.gmm_prec + ((size_t)c * D + r) * D;
            const double* mu = M.gmm_mean + (size_t)c * D;
            double s = 0;
            for (int k = 0; k < D; ++k) s += Pm[k] * (S.aa[k] - mu[k]);
            S.ycomp[c * Dp + r] = s;
        }
        __syncthreads();
        if (tid < 32) {  // p_c = 1/2 (x-mu)^T Sigma^-1 (x-mu) - consts_log[c]; first minimum wins (strict <)
            double bestp = 1.79769313486231570e308, bestsq = 0;
            int best = 0;
            for (int c = 0; c < C; ++c) {
                double s = 0;
                for (int k = tid; k < D; k += 32) s += (S.aa[k] - M.gmm_mean[(size_t)c * D + k]) * S.ycomp[c * Dp + k];
                s = 0.5 * warp_sum(s);
                const double p = s - M.gmm_clog[c];
                if (p < bestp) {
                    bestp = p;
                    bestsq = s;
                    best = c;
                }
            }
            if (tid == 0) {
                S.iscr[0] = best;
                S.scr[40] = 0.5 * sbp * sbp * (bestsq - M.gmm_clog[best]);
            }
        }
        __syncthreads();
        const int best = S.iscr[0];
        const double hb = 0.5 * sbp * sbp;
        for (int i = tid; i < D * D; i += kLmThreads) {
            const int r = i / D, c = i % D;
            S.Hs[(size_t)(6 + r) * P + 6 + c] += hb * M.gmm_prec[((size_t)best * D + r) * D + c];
        }
        for (int r = tid; r < D; r += kLmThreads) S.gs[6 + r] += hb * S.ycomp[best * Dp + r];
        cost += S.scr[40];
    }
    // ---- shape prior (AvatarOptimizer.cpp:708-723) ----
    if (sbs > 0.0) {
        double sq = 0;
        for (int k = 0; k < K; ++k) sq += w[k] * w[k];
        cost += 0.5 * sbs * sbs * sq;
        for (int k = tid; k < K; k += kLmThreads) {
            S.Hs[(size_t)(3 + 3 * J + k) * P + 3 + 3 * J + k] += sbs * sbs;
            S.gs[3 + 3 * J + k] += sbs * sbs * w[k];
        }
    }
    __syncthreads();
    return cost;
}

// CTA-wide lower Cholesky of W (P x P, row-major, in place); returns false when not positive definite
__device__ bool block_cholesky(double* W, int P, int* flag) {
    const int tid = threadIdx.x;
    if (tid == 0) *flag = 1;
    __syncthreads();
    for (int j = 0; j < P; ++j) {
        const double d = W[(size_t)j * P + j];
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0) *flag = 0;
            break;  // uniform: every thread reads the same d
        }
        const double inv = 1.0 / sqrt(d);
        __syncthreads();
        for (int i = j + tid; i < P; i += kLmThreads) W[(size_t)i * P + j] *= inv;  // W[j][j] becomes sqrt(d)
        __syncthreads();
        // after scaling, W[j][j] = sqrt(d); trailing update of the lower triangle
        const int n = P - j - 1;
        for (int e = tid; e < n * n; e += kLmThreads) {
            const int i = j + 1 + e / n, k = j + 1 + e % n;
            if (k <= i) W[(size_t)i * P + k] -= W[(size_t)i * P + j] * W[(size_t)k * P + j];
        }
        __syncthreads();
    }
    __syncthreads();
    return *flag != 0;
}

// solve L L^T x = b (b overwritten) with one warp
__device__ void warp_chol_solve(const double* L, int P, double* b) {
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < P; ++i) {
        double s = 0;
        for (int k = lane; k < i; k += 32) s += L[(size_t)i * P + k] * b[k];
        s = warp_sum(s);
        if (lane == 0) b[i] = (b[i] - s) / L[(size_t)i * P + i];
        __syncwarp();
    }
    for (int i = P - 1; i >= 0; --i) {
        double s = 0;
        for (int k = i + 1 + lane; k < P; k += 32) s += L[(size_t)k * P + i] * b[k];
        s = warp_sum(s);
        if (lane == 0) b[i] = (b[i] - s) / L[(size_t)i * P + i];
        __syncwarp();
    }
}

template <typename AccT>
__global__ void __launch_bounds__(kLmThreads, 1)
lm_fit_kernel(DevModel M, DevParts Pt, LmArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int P = M.P, nx = M.nx, J = M.J, K = M.K, V = M.V;
    LmSmem S = carve_lm(smem_raw, M);
    double* W = reinterpret_cast<double*>(S.A);  // Cholesky workspace aliases the Jacobian tile
    double* Hcur = a.Hcur + (size_t)f * P * P;

    for (int i = tid; i < nx; i += kLmThreads) S.xs[i] = a.x[(size_t)f * nx + i];
    // matched vertices grouped by column group, ascending inside a group; correspondences count
    const int* cnt = a.cnt + (size_t)f * V;
    {
        const int lane = tid & 31, wid = tid >> 5, nw = kLmThreads >> 5;
        int base = 0, ncorr = 0;
        if (tid == 0) S.gcount[0] = 0;
        for (int g = 0; g < Pt.numGroups; ++g) {
            const int b0 = Pt.gvstart[g], b1 = Pt.gvstart[g + 1];
            for (int i0 = b0; i0 < b1; i0 += kLmThreads) {
                const int i = i0 + tid;
                int v = 0, cv = 0;
                if (i < b1) {
                    v = Pt.gorder[i];
                    cv = cnt[v];
                }
                ncorr += cv;
                const unsigned bal = __ballot_sync(0xffffffffu, cv > 0);
                const int wpre = __popc(bal & ((1u << lane) - 1));
                if (lane == 0) S.iscr[wid] = __popc(bal);
                __syncthreads();
                int woff = 0, tot = 0;
                for (int q = 0; q < nw; ++q) {
                    if (q < wid) woff += S.iscr[q];
                    tot += S.iscr[q];
                }
                if (cv > 0) S.mlist[base + woff + wpre] = (unsigned short)v;
                base += tot;
                __syncthreads();
            }
            if (tid == 0) S.gcount[g + 1] = base;
        }
        ncorr = warp_sum_i(ncorr);
        if (lane == 0) S.iscr[32 + wid] = ncorr;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < nw; ++q) t += S.iscr[32 + q];
            S.iscr[48] = t;
            S.iscr[49] = base;
        }
        __syncthreads();
    }
    const int ncorr = S.iscr[48], nmatched = S.iscr[49];
    // sum |d|^2 over the frame's fixed 256-point blocks, in block order
    double Qsum = 0;
    {
        const int qb0 = a.frame_qblock[f], qb1 = a.frame_qblock[f + 1];
        double part = 0;
        for (int i = qb0 + tid; i < qb1; i += kLmThreads) part += a.qpart[i];
        Qsum = block_sum(part, S.scr);
    }
    // scaledBeta{Pose,Shape} = beta * sqrt(#correspondences) / 15 (AvatarOptimizer.cpp:1457-1458)
    const double sbp = a.beta_pose * sqrt((double)ncorr) / 15.0;
    const double sbs = a.beta_shape * sqrt((double)ncorr) / 15.0;

    double cost = lm_evaluate<AccT>(M, Pt, a, S, S.xs, f, Qsum, sbp, sbs);
    for (int i = tid; i < P * P; i += kLmThreads) Hcur[i] = S.Hs[i];
    for (int i = tid; i < P; i += kLmThreads) S.gcur[i] = S.gs[i];
    __syncthreads();
    const double initial_cost = cost;
    if (a.dump_cost) {  // avb_debug_evaluate
        if (tid == 0) a.dump_cost[f] = cost;
        for (int i = tid; i < P; i += kLmThreads) a.dump_grad[(size_t)f * P + i] = S.gcur[i];
        for (int i = tid; i < P * P; i += kLmThreads) a.dump_H[(size_t)f * P * P + i] = S.Hs[i];
    }

    // ---- Levenberg-Marquardt (Ceres-1.14-style trust region, see DESIGN.md "solver") ----
    double radius = 1e4, decrease = 2.0;
    int iters = 0, accepted = 0;
    bool done = false;
    {
        double m = 0;
        for (int i = 0; i < P; ++i) m = fmax(m, fabs(S.gcur[i]));
        done = !(m > 1e-10) || ncorr == 0 || !isfinite(cost);
    }
    for (int it = 0; it < a.max_iters && !done; ++it) {
        ++iters;
        // W = Hcur + D,  D_jj = clamp(s^2 h_jj, 1e-6, 1e32) / (s^2 radius),  s = 1 / (1 + sqrt(h_jj))
        for (int i = tid; i < P * P; i += kLmThreads) {
            double h = Hcur[i];
            const int r = i / P, c = i % P;
            if (r == c) {
                const double s = 1.0 / (1.0 + sqrt(h));
                const double d = fmin(fmax(s * s * h, 1e-6), 1e32);
                h += d / (s * s * radius);
            }
            W[i] = h;
        }
        __syncthreads();
        bool ok = block_cholesky(W, P, &S.iscr[50]);
        double model_change = 0.0;
        if (ok) {
            for (int i = tid; i < P; i += kLmThreads) S.delta[i] = -S.gcur[i];
            __syncthreads();
            if (tid < 32) warp_chol_solve(W, P, S.delta);
            __syncthreads();
            // model_cost_change = -delta^T (g + 1/2 H delta)
            double part = 0;
            for (int r = tid; r < P; r += kLmThreads) {
                double s = 0;
                for (int c = 0; c < P; ++c) s += Hcur[(size_t)r * P + c] * S.delta[c];
                part -= S.delta[r] * (S.gcur[r] + 0.5 * s);
            }
            model_change = block_sum(part, S.scr);
            ok = model_change > 0.0 && isfinite(model_change);
        }
        bool acc = false;
        if (ok) {
            // retraction: p, w additive; q <- dq (x) q, |delta| is the half angle (AvatarOptimizer.cpp:123-143)
            for (int i = tid; i < 3; i += kLmThreads) S.xt[i] = S.xs[i] + S.delta[i];
            for (int k = tid; k < K; k += kLmThreads) S.xt[3 + 4 * J + k] = S.xs[3 + 4 * J + k] + S.delta[3 + 3 * J + k];
            for (int j = tid; j < J; j += kLmThreads) {
                const double* d = S.delta + 3 + 3 * j;
                const double* q = S.xs + 3 + 4 * j;
                double* o = S.xt + 3 + 4 * j;
                const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                if (nd > 0.0) {
                    const double sdd = sin(nd) / nd;
                    const double ax = sdd * d[0], ay = sdd * d[1], az = sdd * d[2], aw = cos(nd);
                    o[0] = aw * q[0] + ax * q[3] + ay * q[2] - az * q[1];
                    o[1] = aw * q[1] + ay * q[3] + az * q[0] - ax * q[2];
                    o[2] = aw * q[2] + az * q[3] + ax * q[1] - ay * q[0];
                    o[3] = aw * q[3] - ax * q[0] - ay * q[1] - az * q[2];
                } else {
                    o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
                }
            }
            __syncthreads();
            const double cost_t = lm_evaluate<AccT>(M, Pt, a, S, S.xt, f, Qsum, sbp, sbs);
            const double rho = (cost - cost_t) / model_change;
            if (isfinite(cost_t) && rho > 1e-3) {
                acc = true;
                ++accepted;
                double dn = 0, xn = 0;
                for (int i = 0; i < nx; ++i) {
                    const double dd = S.xt[i] - S.xs[i];
                    dn += dd * dd;
                    xn += S.xs[i] * S.xs[i];
                }
                const double cost_change = cost - cost_t;
                const double cost_old = cost;
                __syncthreads();
                for (int i = tid; i < nx; i += kLmThreads) S.xs[i] = S.xt[i];
                for (int i = tid; i < P * P; i += kLmThreads) Hcur[i] = S.Hs[i];
                for (int i = tid; i < P; i += kLmThreads) S.gcur[i] = S.gs[i];
                __syncthreads();
                cost = cost_t;
                radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3.0)));
                decrease = 2.0;
                if (sqrt(dn) <= 1e-8 * (sqrt(xn) + 1e-8)) done = true;
                if (fabs(cost_change) <= a.function_tolerance * cost_old) done = true;
                double m = 0;
                for (int i = 0; i < P; ++i) m = fmax(m, fabs(S.gcur[i]));
                if (!(m > 1e-10)) done = true;
            }
        }
        if (!acc) {
            radius /= decrease;
            decrease *= 2.0;
            if (radius < 1e-32) done = true;
        }
        if (a.trace) {
            const int slot = min(it, a.trace_cap - 1);
            for (int i = tid; i < nx; i += kLmThreads) a.trace[((size_t)f * a.trace_cap + slot) * nx + i] = S.xs[i];
        }
        __syncthreads();
    }
    for (int i = tid; i < nx; i += kLmThreads) a.x[(size_t)f * nx + i] = S.xs[i];
    if (tid == 0) {
        FrameStats& st = a.stats[f];
        st.num_correspondences = ncorr;
        st.num_matched_vertices = nmatched;
        st.iterations = iters;
        st.accepted_steps = accepted;
        st.initial_cost = initial_cost;
        st.final_cost = cost;
        st.status = (a.range_flag[f] || !isfinite(cost)) ? 4 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
size_t pose_smem_bytes(int V, int J, int K) {
    const int nx = 3 + 4 * J + K;
    return (size_t)(((nx + 1) & ~1) + tables_doubles(J, K, false)) * 8 + 64 * 4 + (size_t)V + 64;
}

cudaError_t launch_pose_visibility(const DevModel& M, const DevParts& Pt, const PoseArgs& a, int batch, cudaStream_t st) {
    const size_t smem = pose_smem_bytes(M.V, M.J, M.K);
    pose_visibility_kernel<<<batch, 256, smem, st>>>(M, Pt, a);
    return cudaGetLastError();
}

cudaError_t launch_nn(const DevParts& Pt, const NNArgs& a, int num_chunks, cudaStream_t st) {
    const size_t smem = (((size_t)a.V * 24 + 15) & ~(size_t)15) + 128;
    {   // per device/context attribute: set on every launch (cheap host call)
        cudaError_t e = cudaFuncSetAttribute(nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (num_chunks > 0) nn_kernel<<<num_chunks, 512, smem, st>>>(Pt, a);
    return cudaGetLastError();
}

cudaError_t launch_lm(const DevModel& M, const DevParts& Pt, const LmArgs& a, int batch, bool acc64, cudaStream_t st) {
    const size_t smem = lm_smem_bytes(M.V, M.J, M.K, M.gmmC);
    {
        cudaError_t e = acc64 ? cudaFuncSetAttribute(lm_fit_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(lm_fit_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (acc64)
        lm_fit_kernel<double><<<batch, kLmThreads, smem, st>>>(M, Pt, a);
    else
        lm_fit_kernel<float><<<batch, kLmThreads, smem, st>>>(M, Pt, a);
    return cudaGetLastError();
}

}  // namespace avb
