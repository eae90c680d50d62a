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
#include "avb_tables.cuh"

#include <math.h>

#include <algorithm>

namespace avb {

// ---------------------------------------------------------------------------------------------
// K1: forward SMPL + visibility + per-part compaction.  grid = batch, block = 512 (two CTAs per SM: a lane of <= 296 frames is one wave)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
pose_visibility_kernel(DevModel M, DevParts Pt, PoseArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = a.slices > 1 ? a.slices : 1;   // > 1 only for forward-only launches (small batches: most SMs would idle)
    const int f = blockIdx.x / S, slice = blockIdx.x - f * S, tid = threadIdx.x, nt = blockDim.x;
    const int V = M.V, J = M.J, K = M.K;
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* tb = xs + ((M.nx + 1) & ~1);
    Tables T = carve_tables(tb, J, K, false);
    int* iscr = reinterpret_cast<int*>(tb + tables_doubles(J, K, false));  // [64]
    uint8_t* vis = reinterpret_cast<uint8_t*>(iscr + 64);                    // [V]

    for (int i = tid; i < M.nx; i += nt) xs[i] = a.x[(size_t)f * M.nx + i];
    __syncthreads();
    if (a.do_lbs) build_tables(M, xs, T, false);

    double* cloud = a.cloud + (size_t)f * 3 * V;
    const double* w = xs + 3 + 4 * J;
    for (int v = a.do_lbs ? slice * nt + tid : V; v < V; v += nt * S) {
        // shape blend (Avatar.cpp:26) then LBS with the assignedJoints weights (AvatarOptimizer.cpp:507-514)
        const float* sd = M.sdT + v;   // component-major: consecutive threads read consecutive addresses
        double v0[3];
        if (K == 10) {   // the SMPL shape space: all thirty key-cloud loads of the vertex in flight (they are L2 hits: latency, not bytes)
            float t[30];
#pragma unroll
            for (int i = 0; i < 30; ++i) t[i] = __ldg(sd + (size_t)i * V);
            const double b0 = M.vt[3 * (size_t)v], b1 = M.vt[3 * (size_t)v + 1], b2 = M.vt[3 * (size_t)v + 2];
            const double bb[3] = {b0, b1, b2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 10; ++k) s += (double)t[c * 10 + k] * w[k];   // same order as the generic loop
                v0[c] = bb[c] + s;
            }
        } else {
            for (int c = 0; c < 3; ++c) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)sd[(size_t)(c * K + k) * V] * w[k];
                v0[c] = M.vt[3 * (size_t)v + c] + s;
            }
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
    if (a.joint_pos && slice == 0 && a.do_lbs)
        for (int i = tid; i < 3 * J; i += nt) a.joint_pos[(size_t)f * 3 * J + i] = T.pos[i];
    if (a.joint_trans && slice == 0 && a.do_lbs)  // 3x4 column-major per joint: [G_j | tau_j] (Avatar.cpp:59-64)
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
#pragma unroll 4
        for (int t = tid; t < M.F; t += nt) {   // unrolled: the index -> position gathers of four faces in flight
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
    if (a.zero_cnt) {   // the accumulators of the correspondence search that follows
        int* zc = a.zero_cnt + (size_t)f * V;
        unsigned long long* zs = a.zero_sum + (size_t)f * 3 * V;
        for (int v = tid; v < V; v += nt) zc[v] = 0;
        for (int v = tid; v < 3 * V; v += nt) zs[v] = 0ull;
        if (tid == 0) a.zero_range[f] = 0;
    }

    // ---- compaction per part, ascending vertex id inside a part (AvatarOptimizer.cpp:860-880) ----
    int* pv_idx = a.pv_idx + (size_t)f * V;
    double* pv_xyz = a.pv_xyz + (size_t)f * a.pv_stride;
    int* pv_start = a.pv_start + (size_t)f * (Pt.numParts + 1);
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int base = 0;
    float rmax = 0.f;   // a NaN coordinate makes fmaxf skip it; nn_kernel tests its own bound for finiteness
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
                const double c0 = cloud[3 * (size_t)v], c1 = cloud[3 * (size_t)v + 1], c2 = cloud[3 * (size_t)v + 2];
                pv_xyz[3 * (size_t)pos] = c0;
                pv_xyz[3 * (size_t)pos + 1] = c1;
                pv_xyz[3 * (size_t)pos + 2] = c2;
                if (a.pv_f32) {   // relative to the root translation: |m - p| stays within the body's extent
                    const float f0 = (float)(c0 - xs[0]), f1 = (float)(c1 - xs[1]), f2 = (float)(c2 - xs[2]);
                    a.pv_f32[(size_t)f * V + pos] = make_float4(f0, f1, f2, 0.f);
                    rmax = fmaxf(rmax, fmaxf(fabsf(f0), fmaxf(fabsf(f1), fabsf(f2))));
                }
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
    if (a.pv_f32) {   // frame maximum of |float coordinate|
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        __syncthreads();
        float* fscr = reinterpret_cast<float*>(iscr);
        if (lane == 0) fscr[wid] = rmax;
        __syncthreads();
        if (tid == 0) {
            float r = 0.f;
            for (int q = 0; q < nw; ++q) r = fmaxf(r, fscr[q]);
            a.pv_rmax[f] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: exact fp64 1-NN, data -> visible model vertices of the same part.  block = 512
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared memory holds the first a.stage_cap compacted vertices (host: what lets TWO CTAs share an SM; back-face culling
// leaves about half of the model visible, so a frame's whole visible cloud normally fits); a part whose range ends beyond
// the staged prefix is scanned from global memory (L2) instead -- same arithmetic, same order.
constexpr int kNNPer = 4;               // points per thread between two CTA barriers
constexpr int kNNPass = 512 * kNNPer;

template <bool F32DATA>   // the data cloud is the float upload (a.data_f32) / the doubles (a.data)
__global__ void __launch_bounds__(512, 2)
nn_kernel(DevParts Pt, NNArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int s_start[kMaxParts + 1];
    __shared__ double s_wsum[16 * kNNPer];
    __shared__ int s_hist[kMaxParts + 1], s_off[kMaxParts + 1];
    __shared__ unsigned short s_perm[kNNPass];
    __shared__ double s_q2[kNNPass];
    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    const int f = a.chunk_frame[c];
    const long long begin = a.chunk_begin[c];
    const int count = a.chunk_count[c];
    double* mstage = reinterpret_cast<double*>(smem_raw);

    const int* pv_start = a.pv_start + (size_t)f * (Pt.numParts + 1);
    if (tid <= Pt.numParts) s_start[tid] = pv_start[tid];
    const bool fast = a.pv_f32 != nullptr;   // stage the float copy (16 B per vertex) for the pre-scan instead of the doubles
    const int nvis = min(pv_start[Pt.numParts], a.stage_cap);
    // one TMA bulk copy of the (staged prefix of the) frame's compacted model cloud (nvis * 24 B or 16 B, rounded up to 16 B)
    const uint32_t bytes = (uint32_t)(((size_t)nvis * (fast ? 16 : 24) + 15) & ~(size_t)15);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && bytes > 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes)
                     : "memory");
        const void* src = fast ? static_cast<const void*>(a.pv_f32 + (size_t)f * a.V) : static_cast<const void*>(a.pv_xyz + (size_t)f * a.pv_stride);
        uint32_t off = 0;
        while (off < bytes) {  // <= 64 KiB per bulk copy
            const uint32_t n = min(bytes - off, 65536u);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(reinterpret_cast<unsigned char*>(mstage) + off)),
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
    // passes of kNNPass = 2048 points = eight deterministic |d|^2 blocks of kQBlock (256) points, four points per thread.
    // Inside a pass the points are dealt to the threads SORTED BY PART LABEL (a counting sort in shared memory): image-ordered
    // clouds put two or more labels into most warps, and a warp whose lanes scan different parts runs those scans one after
    // the other.  Thread t takes the sorted positions t, t + 512, ... so every warp works through four different label runs
    // between two CTA barriers: parts differ tenfold in size, and with one point per thread a fifth of the kernel was spent
    // waiting for the warp that drew the largest part.  Which thread handles which point does not matter for the result: the
    // index goes back to the point's own slot, the per-vertex sums are integer atomics, and |d|^2 returns to the point's
    // original position before the block sums.
    for (int base = 0; base < count; base += kNNPass) {
        if (tid <= kMaxParts) s_hist[tid] = 0;
        __syncthreads();
        int mylab[kNNPer], slot[kNNPer];
#pragma unroll
        for (int j = 0; j < kNNPer; ++j) {
            const int li0 = base + j * 512 + tid;
            mylab[j] = kMaxParts;   // bucket of "nothing to scan"
            if (li0 < count) {
                const int label = a.labels[begin + li0];
                if (label >= 0 && label < Pt.numParts) mylab[j] = label;
                else atomicOr(a.range_flag + f, 2);   // out of range is UB in the reference (:1279)
            }
            slot[j] = atomicAdd(&s_hist[mylab[j]], 1);
        }
        __syncthreads();
        if (tid == 0) {   // exclusive prefix over the label buckets
            int run = 0;
            for (int l = 0; l <= kMaxParts; ++l) {
                s_off[l] = run;
                run += s_hist[l];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kNNPer; ++j) s_perm[s_off[mylab[j]] + slot[j]] = (unsigned short)(j * 512 + tid);
        __syncthreads();
#pragma unroll 1
        for (int j = 0; j < kNNPer; ++j) {
        const int src = s_perm[j * 512 + tid], li = base + src;
        double q2 = 0.0;
        if (li < count) {
            const long long gi = begin + li;
            double q0, q1, qz;
            if (F32DATA) {   // float upload: widened here (exact), no separate pass over the cloud
                q0 = (double)a.data_f32[3 * gi];
                q1 = (double)a.data_f32[3 * gi + 1];
                qz = (double)a.data_f32[3 * gi + 2];
            } else {
                q0 = a.data[3 * gi];
                q1 = a.data[3 * gi + 1];
                qz = a.data[3 * gi + 2];
            }
            const int label = a.labels[gi];
            const bool lab_ok = label >= 0 && label < Pt.numParts;
            const int s = lab_ok ? s_start[label] : 0, e = lab_ok ? s_start[label + 1] : 0;
            double best = 1.79769313486231570e308;
            int bi = -1;
            bool need_exact = e > s;
            if (fast && e > s && e <= a.stage_cap) {
                // ---- fp32 pre-scan (exact RESULT, inexact SEARCH): best and second-best float distance to the staged float
                //      copy (coordinates relative to the root translation p).  If the two are separated by more than the
                //      rigorous error bound of the float pipeline, the best candidate IS the exact nearest neighbour and no
                //      fp64 distance is ever evaluated; otherwise (near-ties, non-finite input) the exact scan below runs.
                //      Error of one float coordinate difference: conversions of q - p and m - p (2^-24 relative each) and the
                //      subtraction (2^-24 of the result) <= 2^-24 (|q-p| + |m-p| + |diff|) <= e = 2^-22 R, R >= every
                //      |coordinate|.  Hence |d~ - d| <= 2 sqrt(3 d) e + 3 e^2 + 2^-21 d~ =: E(d~) (products and sums: 3 ulp).
                const float4* fst = reinterpret_cast<const float4*>(mstage);
                const double* px = a.x + (size_t)f * a.nx;
                const float qx = (float)(q0 - px[0]), qy = (float)(q1 - px[1]), qw = (float)(qz - px[2]);
                const float R = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fmaxf(fabsf(qw), a.pv_rmax[f]));
                const float eb = 2.4e-7f * R;   // 2^-22 R, rounded up
                float b1 = __int_as_float(0x7f800000), b2 = b1;
                int i1 = -1;
#pragma unroll 4
                for (int k = s; k < e; ++k) {
                    const float4 m = fst[k];
                    const float dx = qx - m.x, dy = qy - m.y, dz = qw - m.z;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    b2 = fminf(b2, fmaxf(d, b1));
                    if (d < b1) i1 = k;
                    b1 = fminf(b1, d);
                }
                const float E1 = (3.4642f * sqrtf(b1) * eb + 3.f * eb * eb + 4.8e-7f * b1) * 1.01f;
                const float E2 = (3.4642f * sqrtf(b2) * eb + 3.f * eb * eb + 4.8e-7f * b2) * 1.01f;
                // x - E(x) is increasing for x > 3 e^2, so every other candidate's exact distance exceeds b2 - E(b2)
                if (i1 >= 0 && b2 > 4.f * eb * eb && b2 - E2 > b1 + E1) {   // false for NaN / inf anywhere
                    bi = i1;
                    need_exact = false;
                }
            }
            const double* mxyz = (!fast && e <= a.stage_cap) ? mstage : a.pv_xyz + (size_t)f * a.pv_stride;
#pragma unroll 4
            for (int k = s; k < (need_exact ? e : s); ++k) {
                // nanoflann L2_Simple: sum_c (a_c - b_c)^2, c = 0,1,2 (nanoflann.hpp:423-445); strict '<'
                const double d0 = q0 - mxyz[3 * k], d1 = q1 - mxyz[3 * k + 1], d2 = qz - mxyz[3 * k + 2];
                // ((d0^2 + d1^2) + d2^2) with every product and sum rounded: the reference is built without -march
                // (CMakeLists.txt:37), so `result += diff * diff` is never contracted into an FMA
                const double dist = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
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
        s_q2[src] = q2;
        }   // points of this thread
        __syncthreads();
        // deterministic per-256-point partials of sum |d|^2, in the points' own order
#pragma unroll
        for (int j = 0; j < kNNPer; ++j) {
            const double v = warp_sum(s_q2[j * 512 + tid]);
            if (lane == 0) s_wsum[j * 16 + wid] = v;
        }
        __syncthreads();
        if (tid < 2 * kNNPer) {
            const int blk = (base / kQBlock) + tid;  // block index inside this chunk
            if (blk * kQBlock < count) {
                double sacc = 0;
                for (int q = 0; q < 8; ++q) sacc += s_wsum[(tid >> 1) * 16 + (tid & 1) * 8 + q];
                a.qpart[a.chunk_qblock[c] + blk] = sacc;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// float -> double widening of an uploaded cloud (avb_upload_batch_f32): depth cameras deliver float points
// (CameraIntrin::to3D, Calibration.cpp:68-95), so shipping them as floats halves the PCIe bytes and the widened
// doubles are bit-identical to what the reference's Eigen::Vector3f -> double cast produces on the host (demo.cpp:241)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
widen_points_kernel(const float4* __restrict__ in, double* __restrict__ out, long long n4, const float* __restrict__ tail_in, int tail) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(in + i);
        double2* o = reinterpret_cast<double2*>(out + 4 * i);
        o[0] = make_double2((double)v.x, (double)v.y);
        o[1] = make_double2((double)v.z, (double)v.w);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) out[4 * n4 + threadIdx.x] = (double)tail_in[threadIdx.x];
}

cudaError_t launch_widen_points(const float* in, double* out, long long n, int num_sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const long long n4 = n / 4;
    const int tail = (int)(n - 4 * n4);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)num_sms * 8));
    widen_points_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), out, n4, in + 4 * n4, tail);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
size_t pose_smem_bytes(int V, int J, int K) {
    const int nx = 3 + 4 * J + K;
    return (size_t)(((nx + 1) & ~1) + tables_doubles(J, K, false)) * 8 + 64 * 4 + (size_t)V + 64;
}

size_t nn_smem_bytes(int stage_cap, bool f32) { return (((size_t)stage_cap * (f32 ? 16 : 24) + 15) & ~(size_t)15) + 128; }

cudaError_t launch_pose_visibility(const DevModel& M, const DevParts& Pt, const PoseArgs& a, int batch, cudaStream_t st) {
    const size_t smem = pose_smem_bytes(M.V, M.J, M.K);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pose_visibility_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int S = (a.slices > 1 && !a.do_visibility) ? a.slices : 1;
    PoseArgs b = a;
    b.slices = S;
    pose_visibility_kernel<<<batch * S, 512, smem, st>>>(M, Pt, b);
    return cudaGetLastError();
}

cudaError_t launch_nn(const DevParts& Pt, const NNArgs& a, int num_chunks, cudaStream_t st) {
    const size_t smem = nn_smem_bytes(a.stage_cap, a.pv_f32 != nullptr);
    {   // per device/context attribute: set on every launch (cheap host call)
        cudaError_t e = a.data_f32 ? cudaFuncSetAttribute(nn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                   : cudaFuncSetAttribute(nn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (num_chunks > 0) {
        if (a.data_f32) nn_kernel<true><<<num_chunks, 512, smem, st>>>(Pt, a);
        else nn_kernel<false><<<num_chunks, 512, smem, st>>>(Pt, a);
    }
    return cudaGetLastError();
}

}  // namespace avb
