// avb_lm.cu -- the inner solve of one ICP iteration (AvatarOptimizer.cpp:1398-1486) as grid-sized kernels.
//
//   lm_prep_kernel   grid = frames            matched-vertex lists per Jacobian column group, chunk list,
//                                             #correspondences, sum |d|^2, LM state, joint tables at x.
//   lm_rows_kernel   grid = 256-vertex blocks x frames   one thread per matched vertex: position, residual
//                                             statistics and the analytic Jacobian (AvatarOptimizer.cpp:505-582 in
//                                             closed form) as a compact fp32 record (SoA) in HBM; cost partial.
//   lm_gram_kernel   grid = chunks x frames   Gram matrix of the records of up to 256 matched vertices (fp64 DMMA from
//                                             an fp32 tile), turned into the chunk's deterministic partial of
//                                             J^T J / J^T r.
//   lm_flow_kernel<true> (default)            persistent data-flow kernel: fused record + Gram tasks (fused_body: the
//                                             records go straight into a swizzled bf16 operand tile, J^T J on tcgen05
//                                             with a TMEM accumulator, J^T r and cost in fp64) and the solves.
//   lm_flow_kernel<false>                     the same data flow over rows / gram (fp64 DMMA) / solve tasks.
//   lm_solve_kernel  grid = frames            partial reduction in chunk order, priors (:661-692, :708-723),
//                                             Levenberg-Marquardt step control, damped Cholesky solve, retraction
//                                             (:123-143), joint tables of the next trial point.
//
// One evaluation = rows + gram + solve; 1 + maxItersPerICP evaluations per ICP iteration, enqueued back to back on
// one stream with no host round trip (frames that converged early skip their CTAs).
//
// Numerics: geometry, residuals, cost, gradient and the linear algebra are fp64.  Jacobian records are stored as
// fp32 (they only scale the residual in J^T r, so their 6e-8 rounding moves the fixed point by ~1e-11; the residual
// itself is carried as an fp32 hi/lo pair) and every product and sum of J^T J / J^T r is fp64.
#include "avb_device.cuh"
#include "avb_kernels.h"
#include "avb_tables.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace avb {

constexpr int kJacThreads = 256;
constexpr int kSolveThreads = 256;
constexpr unsigned short kNoVertex = 0xFFFFu;   // gap slot in mlist

// Buffers that one task writes and a later task (possibly on another SM) reads inside the same lm_flow_kernel launch
// are read through L2 (ld.global.cg): L1 is not coherent across SMs.
__device__ __forceinline__ double ldg2(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int ldg2(const int* p) { return __ldcg(p); }
__device__ __forceinline__ LmState load_state(const LmState* p) {
    static_assert(sizeof(LmState) % 16 == 0, "LmState is copied as 16-byte words");
    LmState st;
    const longlong2* src = reinterpret_cast<const longlong2*>(p);
    longlong2* dst = reinterpret_cast<longlong2*>(&st);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(LmState) / 16); ++i) dst[i] = __ldcg(src + i);
    return st;
}

// ---------------------------------------------------------------------------------------------
// work queue of lm_flow_kernel (tasks are pushed only when their inputs are complete, so no task ever waits)
// ---------------------------------------------------------------------------------------------
enum { kTaskRows = 0, kTaskGram = 1, kTaskFused = 2, kTaskPrior = 3 };
__device__ __forceinline__ unsigned make_task(int type, int f, int idx) {
    return ((unsigned)type << 30) | ((unsigned)f << 12) | (unsigned)idx;
}
// one thread: append tasks (type, f, 0..n-1); the caller has fenced the data the tasks will read
__device__ void flow_push(const FlowQueue& q, int type, int f, int n) {
    const unsigned t0 = atomicAdd(&q.ctrl[1], (unsigned)n);
    for (int i = 0; i < n; ++i) {
        const unsigned t = t0 + (unsigned)i;
        const unsigned long long v = ((unsigned long long)t << 32) | make_task(type, f, i);
        *reinterpret_cast<volatile unsigned long long*>(q.slots + (t & q.cap_mask)) = v;
    }
}
// one thread: next task, or -1 when every frame of the launch has finished (or the watchdog fired).  A ticket is
// taken unconditionally (one atomicAdd: a compare-and-swap loop collapses under ~300 contending CTAs) and the CTA
// waits for the ticket's slot; tickets taken past the last push are reconciled by the last CTA to leave the kernel.
__device__ int flow_pop(const FlowQueue& q) {
    volatile unsigned* ctrl = q.ctrl;
    const unsigned ticket = atomicAdd(&q.ctrl[0], 1u);
    volatile unsigned long long* slot = q.slots + (ticket & q.cap_mask);
    const long long t_begin = clock64();
    unsigned long long v = *slot;
    while ((unsigned)(v >> 32) != ticket) {
        if (ctrl[2] == 0u || ctrl[3] != 0u) return -1;
        __nanosleep(128);
        if (clock64() - t_begin > (8ll << 30)) {   // ~4 s without work: report instead of hanging the device
            atomicExch(&q.ctrl[3], 1u);
            return -1;
        }
        v = *slot;
    }
    __threadfence();
    return (int)(unsigned)(v & 0xFFFFFFFFull);
}
// one thread, on leaving the kernel: the last CTA out sets head = tail for the next launch
__device__ void flow_leave(const FlowQueue& q) {
    __threadfence();
    if (atomicAdd(&q.ctrl[4], 1u) == gridDim.x - 1) {
        q.ctrl[4] = 0u;
        q.ctrl[0] = *reinterpret_cast<volatile unsigned*>(&q.ctrl[1]);
        __threadfence();
    }
}

// one thread: the tasks that open an evaluation of frame f (its state and trial point are written and fenced).  The frame's
// solve runs when gram_left[f] reaches zero: one count per Gram / fused task (cost-only evaluation of the two-task schedule:
// one for "all record tasks done") plus one for the prior task, if the prior runs as a task.
__device__ void flow_push_eval(const LmBuf& a, int f, int nslots, int nchunks, bool cost_only, bool want_prior) {
    const int extra = want_prior ? 1 : 0;
    if (a.tensor || a.fused) {
        atomicExch(&a.q.gram_left[f], nchunks + extra);
        __threadfence();
        flow_push(a.q, kTaskFused, f, nchunks);
    } else {
        const int nrb = (nslots + 255) >> 8;
        atomicExch(&a.q.rows_left[f], nrb);
        atomicExch(&a.q.gram_left[f], (cost_only ? 1 : nchunks) + extra);
        __threadfence();
        flow_push(a.q, kTaskRows, f, nrb);
    }
    if (want_prior) flow_push(a.q, kTaskPrior, f, 1);
}

// avb_set_profiling: thread 0 adds the nanoseconds since *t_prev to phase counter `cls` (prof[4..15]: sub-phases)
__device__ __forceinline__ void phase_lap(const FlowQueue& q, int cls, unsigned long long& t_prev) {
    if (q.prof && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicAdd(q.prof + cls, t - t_prev);
        t_prev = t;
    }
    if (q.prof) __syncwarp();   // thread 0 rejoins its warp here: a warp left diverged runs its shuffle / __syncwarp code on the slow path
}
__device__ __forceinline__ unsigned long long phase_begin(const FlowQueue& q) {
    unsigned long long t = 0;
    if (q.prof && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__host__ __device__ inline int tab_doubles(int J, int K) { return J * (15 + 3 * K); }

// columns of a group's compact Jacobian: [ p(3) | 3 per group joint | K shape ]
__host__ __device__ inline int group_L(int nj, int K) { return 3 + 3 * nj + K; }

// ---------------------------------------------------------------------------------------------
// lm_prep_kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
lm_prep_kernel(DevModel M, DevParts Pt, LmBuf a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int V = M.V, J = M.J, K = M.K, nx = M.nx;
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* tb = xs + ((nx + 1) & ~1);
    double* scr = tb + tables_doubles(J, K, true);
    int* iscr = reinterpret_cast<int*>(scr + 64);
    Tables T = carve_tables(tb, J, K, true);

    for (int i = tid; i < nx; i += nt) {
        const double v = a.x[(size_t)f * nx + i];
        xs[i] = v;
        a.xt[(size_t)f * nx + i] = v;
    }
    const int* cnt = a.cnt + (size_t)f * V;
    unsigned short* mlist = a.mlist + (size_t)f * a.rec_rs;
    int4* chunks = a.chunks + (size_t)f * a.maxc;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int base = 0, nmatched = 0, ncorr = 0, nchunks = 0;
    for (int g = 0; g < Pt.numGroups; ++g) {
        const int b0 = Pt.gvstart[g], b1 = Pt.gvstart[g + 1];
        // a group's slots start on a 16-byte boundary of the fp32 record arrays (cp.async / float4 tile loads)
        if (tid < ((4 - (base & 3)) & 3)) mlist[base + tid] = kNoVertex;
        base = (base + 3) & ~3;
        const int gbase = base;
        if (tid == 0) a.gstart[(size_t)f * (kMaxGroups + 1) + g] = base;
        for (int i0 = b0; i0 < b1; i0 += nt) {
            const int i = i0 + tid;
            int v = 0, cv = 0;
            if (i < b1) {
                v = Pt.gorder[i];
                cv = cnt[v];
            }
            ncorr += cv;
            const unsigned bal = __ballot_sync(0xffffffffu, cv > 0);
            const int wpre = __popc(bal & ((1u << lane) - 1));
            if (lane == 0) iscr[wid] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int q = 0; q < nw; ++q) {
                if (q < wid) woff += iscr[q];
                tot += iscr[q];
            }
            if (cv > 0) mlist[base + woff + wpre] = (unsigned short)v;
            base += tot;
            nmatched += tot;
            __syncthreads();
        }
        // chunk list of this group (every thread computes the same numbers; thread 0 writes)
        const int c0 = min(nchunks, a.maxc);
        for (int s = gbase; s < base; s += a.chunk_verts) {
            if (tid == 0 && nchunks < a.maxc) chunks[nchunks] = make_int4(g, s, min(a.chunk_verts, base - s), 0);
            ++nchunks;
        }
        if (tid == 0) a.gruns[(size_t)f * kMaxGroups + g] = make_int2(c0, min(nchunks, a.maxc) - c0);
    }
    if (tid == 0) a.gstart[(size_t)f * (kMaxGroups + 1) + Pt.numGroups] = base;
    ncorr = warp_sum_i(ncorr);
    if (lane == 0) iscr[32 + wid] = ncorr;
    __syncthreads();
    ncorr = 0;
    for (int q = 0; q < nw; ++q) ncorr += iscr[32 + q];
    double part = 0;
    for (int i = a.frame_qblock[f] + tid; i < a.frame_qblock[f + 1]; i += nt) part += a.qpart[i];
    const double Qsum = block_sum(part, scr);

    build_tables(M, xs, T, true);
    double* tab = a.tab + (size_t)f * a.tabD;
    for (int i = tid; i < 9 * J; i += nt) tab[i] = T.G[i];
    for (int i = tid; i < 3 * J; i += nt) {
        tab[9 * J + i] = T.pos[i];
        tab[12 * J + i] = T.tau[i];
    }
    for (int i = tid; i < 3 * J * K; i += nt) tab[15 * J + i] = T.C[i];
    __threadfence();
    __syncthreads();

    if (tid == 0) {
        LmState& st = a.state[f];
        st.cost = 0;
        st.radius = 1e4;   // Ceres initial_trust_region_radius (the value the reference leaves commented at :1319)
        st.decrease = 2.0;
        st.Qsum = Qsum;
        // scaledBeta{Pose,Shape} = beta * sqrt(#correspondences) / 15 (AvatarOptimizer.cpp:1457-1458)
        st.sbp = a.beta_pose * sqrt((double)ncorr) / 15.0;
        st.sbs = a.beta_shape * sqrt((double)ncorr) / 15.0;
        st.initial_cost = 0;
        st.model_change = 0;
        st.done = (ncorr == 0) ? 1 : 0;
        st.iters = 0;
        st.accepted = 0;
        st.ncorr = ncorr;
        st.nmatched = nmatched;
        st.nslots = base;
        st.nchunks = min(nchunks, a.maxc);
        st.evals = 0;
        st.last = 0;
        st.pad0 = st.pad1 = st.pad2 = 0;
        FrameStats& fs = a.stats[f];
        fs.num_correspondences = ncorr;
        fs.num_matched_vertices = nmatched;
        fs.iterations = 0;
        fs.accepted_steps = 0;
        fs.initial_cost = 0;
        fs.final_cost = 0;
        fs.status = a.range_flag[f] ? 4 : 0;
        if (a.q.slots && !st.done) {   // lm_flow_kernel: the frame's first tasks
            __threadfence();
            atomicAdd(&a.q.ctrl[2], 1u);
            flow_push_eval(a, f, base, st.nchunks, false, a.prior_task && st.sbp > 0.0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lm_rows_kernel: one thread per matched vertex -> compact fp32 Jacobian record in HBM
// ---------------------------------------------------------------------------------------------
// Tangent Jacobian in "global-frame rotation" coordinates eta_j = G_parent(j) delta_j:
// block_j = R(-1,parent j) dRot Lq_j = -2 [y_j]x G_parent(j) (AvatarOptimizer.cpp:529-565 in closed form); the
// change of coordinates is undone in lm_solve_kernel.  Record of a vertex with count c (sc = sqrt(c)):
//   [ 2 sc y_j (3 per group joint) | sc | rho_hi(3) | rho_lo(3) | sc S (3 x K, row-major) ],  rho = (c x - sum d)/sc
// stored SoA: field q of slot i at rec[q * rec_rs + i]; slots = matched vertices in group order, every group
// starting on a multiple of four (gap slots hold kNoVertex in mlist and are never read by the Gram kernels).
__host__ __device__ inline int rec_floats(int nj, int K) { return 3 * nj + 3 * K + 7; }

// One matched vertex: its compact Jacobian record (fields at rec[q * RS]: the global SoA of the staged / two-task schedule,
// or a column of the shared-memory tile of the fused task) and its cost term x . (c x - 2 s).
// (Rolled loops over the K shape keys on purpose: a K = 10 specialisation that unrolls them was measured 2 % SLOWER -- 2.7 x
// the code, index arrays demoted to local memory -- tasks run cold, code size is time.)
__device__ __forceinline__ double record_vertex(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, int v, int g,
                                                float* rec, size_t RS, bool cost_only, const double* tab, const double* w) {
    const int J = M.J, K = M.K;
    const double* G = tab;
    const double* pos = tab + 9 * J;
    const double* tau = tab + 12 * J;
    const double* C = tab + 15 * J;
    double costv = 0.0;
    const int nj = Pt.gnj[g];
    const int* gj = Pt.gjoints + g * kMaxJ;
    const float* sd = M.sd + (size_t)v * 3 * K;
    double v0[3];
    if ((K & 1) == 0) {   // even K: the vertex's 3 K floats start on an 8-byte boundary: half as many (8-byte) L2 loads, same order of sums
        const float2* sd2 = reinterpret_cast<const float2*>(sd);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < K; k += 2) {
                const float2 t = sd2[(c * K + k) >> 1];
                s += (double)t.x * w[k];
                s += (double)t.y * w[k + 1];
            }
            v0[c] = M.vt[3 * (size_t)v + c] + s;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)sd[c * K + k] * w[k];
            v0[c] = M.vt[3 * (size_t)v + c] + s;
        }
    }
    const int n = M.sk_n[v];
    double xk[AVB_MAX_ASSIGN_][3], wk[AVB_MAX_ASSIGN_];
    int jk[AVB_MAX_ASSIGN_];
    uint32_t mk[AVB_MAX_ASSIGN_];
    double x[3] = {0, 0, 0}, B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
        if (q < n) {
            const int k = M.sk_j[4 * (size_t)v + q];
            const double wt = M.sk_w[4 * (size_t)v + q];
            const double* Gk = G + 9 * k;
            jk[q] = k;
            wk[q] = wt;
            mk[q] = M.anc_mask[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                xk[q][c] = Gk[3 * c] * v0[0] + Gk[3 * c + 1] * v0[1] + Gk[3 * c + 2] * v0[2] + tau[3 * k + c];
                x[c] += wt * xk[q][c];
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) B[e] += wt * Gk[e];
        } else {
            jk[q] = 0; wk[q] = 0; mk[q] = 0;
            xk[q][0] = xk[q][1] = xk[q][2] = 0;
        }
    }
    const double cn = (double)a.cnt[(size_t)f * M.V + v];
    const double sc = sqrt(cn), s2 = 2.0 * sc;
    // the evaluation that only decides the last accept / reject needs the cost, not the Jacobian records
    for (int gi = 0; gi < (cost_only ? 0 : nj); ++gi) {
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
        rec[(3 * gi) * RS] = (float)((y0 - W * pos[3 * j]) * s2);
        rec[(3 * gi + 1) * RS] = (float)((y1 - W * pos[3 * j + 1]) * s2);
        rec[(3 * gi + 2) * RS] = (float)((y2 - W * pos[3 * j + 2]) * s2);
    }
    // shape: sum_k w_k (G_k (Delta_v - S_k) + H_k) = B Delta_v + sum_k w_k C_k  (AvatarOptimizer.cpp:568-580)
    float* rr = rec + (size_t)(3 * nj) * RS;   // sc | rho_hi | rho_lo
    float* rs = rr + 7 * RS;
    for (int m = 0; m < (cost_only ? 0 : K); ++m) {
        const double d0 = sd[m], d1 = sd[K + m], d2 = sd[2 * K + m];
        double e0 = B[0] * d0 + B[1] * d1 + B[2] * d2;
        double e1 = B[3] * d0 + B[4] * d1 + B[5] * d2;
        double e2 = B[6] * d0 + B[7] * d1 + B[8] * d2;
#pragma unroll
        for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
            if (q < n) {
                const double* Cq = C + (size_t)jk[q] * 3 * K;
                e0 += wk[q] * Cq[m];
                e1 += wk[q] * Cq[K + m];
                e2 += wk[q] * Cq[2 * K + m];
            }
        }
        rs[m * RS] = (float)(e0 * sc);
        rs[(K + m) * RS] = (float)(e1 * sc);
        rs[(2 * K + m) * RS] = (float)(e2 * sc);
    }
    // residual sum of the vertex's correspondences, c x - sum d (AvatarOptimizer.cpp:632-639), split hi/lo
    const unsigned long long* sumv = a.sum + 3 * ((size_t)f * M.V + v);
    if (!cost_only) rr[0] = (float)sc;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double sr = (double)(long long)sumv[c] * kFixInv;
        if (!cost_only) {
            const double rho = (cn * x[c] - sr) / sc;
            const float hi = (float)rho;
            rr[(1 + c) * RS] = hi;
            rr[(4 + c) * RS] = (float)(rho - (double)hi);
        }
        costv += x[c] * (cn * x[c] - 2.0 * sr);  // sum_i |x - d_i|^2 - sum_i |d_i|^2 = x . (c x - 2 s)
    }
    return costv;
}

__device__ void rows_body(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, int blk, int nslots,
                          bool cost_only, unsigned char* smem_raw) {
    const int tid = threadIdx.x;
    const int i = blk * 256 + tid;
    const int J = M.J, K = M.K;
    unsigned long long tp = phase_begin(a.q);
    double* tab = reinterpret_cast<double*>(smem_raw);
    double* w = tab + a.tabD;
    double* scr = w + ((K + 1) & ~1);
    int* gstart = reinterpret_cast<int*>(scr + 32);
    const double* gtab = a.tab + (size_t)f * a.tabD;
    if ((a.tabD & 1) == 0) {   // 16-byte loads, all of a thread's loads in flight (the table comes from L2: latency, not bytes)
        const double2* g2 = reinterpret_cast<const double2*>(gtab);
        double2* t2 = reinterpret_cast<double2*>(tab);
#pragma unroll 4
        for (int q = tid; q < (a.tabD >> 1); q += 256) t2[q] = __ldcg(g2 + q);
    } else {
        for (int q = tid; q < a.tabD; q += 256) tab[q] = ldg2(gtab + q);
    }
    for (int q = tid; q < K; q += 256) w[q] = ldg2(a.xt + (size_t)f * M.nx + 3 + 4 * J + q);
    for (int q = tid; q <= Pt.numGroups; q += 256) gstart[q] = a.gstart[(size_t)f * (kMaxGroups + 1) + q];
    int* s_v = reinterpret_cast<int*>(gstart + kMaxGroups + 2);
    s_v[tid] = (i < nslots) ? (int)a.mlist[(size_t)f * a.rec_rs + i] : (int)kNoVertex;
    __syncthreads();
    phase_lap(a.q, 12, tp);
    double costv = 0.0;
    if (s_v[tid] != (int)kNoVertex) {
        int g = 0;
        while (g + 1 < Pt.numGroups && i >= gstart[g + 1]) ++g;
        float* rec = a.rec + (size_t)f * a.rec_stride * a.rec_rs + i;
        costv = record_vertex(M, Pt, a, f, s_v[tid], g, rec, (size_t)a.rec_rs, cost_only, tab, w);
    }
    const double cs = block_sum(costv, scr);
    if (tid == 0) a.cpart[(size_t)f * a.maxrb + blk] = cs;
    phase_lap(a.q, 13, tp);
}

__global__ void __launch_bounds__(256, 3)
lm_rows_kernel(DevModel M, DevParts Pt, LmBuf a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int f = blockIdx.y;
    const int nslots = a.state[f].nslots;
    if (a.state[f].done || blockIdx.x * 256 >= nslots) return;
    rows_body(M, Pt, a, f, blockIdx.x, nslots, a.state[f].last != 0, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// lm_gram_kernel: J^T J / J^T r of one chunk of Jacobian records through their Gram matrix (fp64 DMMA)
// ---------------------------------------------------------------------------------------------
// A matched vertex contributes three Jacobian rows  sc [ I | -2 [y_j]x ... | S ]  whose products are all bilinear
// in the ONE record row  rec = [ u_j = 2 sc y_j | sc | rho_hi | rho_lo | sc S ]:
//   rot j x rot k   (u_j . u_k) I - u_k u_j^T        rot j x shape m   u_j x (sc S_m)      rot j x rho   u_j x rho
// so J^T J and J^T r follow from the Gram matrix  Gm = sum_v rec_v rec_v^T  (nf x nf, nf = 3 nj + 7 + 3 K) with a
// third of the multiply-adds of the expanded 3-row form and no expansion step.  The Gram matrix is accumulated in
// fp64 on the tensor cores' DMMA path (mma.sync.m8n8k4.f64; tools/ubench/dmma_ubench.cu measures 93-97% of the fp64
// pipe from an fp32 tile, against 65% for an 8x8 SIMT register block) from an fp32 field-major tile T[q][t] that is
// a straight asynchronous copy (cp.async, 16 B) of the chunk's SoA records.  The epilogue turns the Gram matrix into
// the chunk's partial of (J^T J, J^T r) in group-local column order; lm_solve_kernel adds the partials in chunk order.
constexpr int kGramThreads = 256;
constexpr int kGramSlots = 4;     // 2x2-block warp tiles per warp: 28 tiles (14 blocks) over 8 warps

__host__ __device__ inline int num_pairs(int n) { return n * (n + 1) / 2; }
__device__ __forceinline__ void pair_to_blocks(int pair, int n, int& bi, int& bj) {
    bi = 0;
    int rowlen = n;
    while (pair >= rowlen) {
        pair -= rowlen;
        ++bi;
        --rowlen;
    }
    bj = bi + pair;
}
__device__ __forceinline__ int pair_index(int bi, int bj, int n) { return bi * n - ((bi * (bi - 1)) >> 1) + (bj - bi); }

// upper triangle (ra <= rb) of an Lg x Lg matrix enumerated densely: rows r and Lg-1-r share one line of Lg+1 entries
__host__ __device__ inline int tri_count(int Lg) { return ((Lg + 1) >> 1) * (Lg + 1); }
__device__ __forceinline__ bool tri_decode(int idx, int Lg, int& ra, int& rb) {
    const int r = idx / (Lg + 1), c = idx - r * (Lg + 1);
    if (c < Lg - r) {
        ra = r;
        rb = r + c;
        return true;
    }
    ra = Lg - 1 - r;
    rb = ra + (c - (Lg - r));
    return ra != r;   // the middle row of an odd Lg pairs with itself: its second half is empty
}

// Gram matrix in shared memory: 8x8 blocks of the upper block triangle, row-major inside a block
__device__ __forceinline__ double gm(const double* Gs, int n, int p, int q) {
    if (p > q) {
        const int t = p;
        p = q;
        q = t;
    }
    return Gs[pair_index(p >> 3, q >> 3, n) * 64 + ((p & 7) << 3) + (q & 7)];
}

// chunk partial [ upper triangle of J^T J (tri_decode order) | J^T r ] in the group's column order
// [ p(3) | 3 per group joint | K shape ], from the Gram matrix of the records
__device__ void emit_partial(const double* Gs, int n, int nj, int K, double* part, int tid, int nt) {
    const int Lg = 3 + 3 * nj + K, nH = tri_count(Lg);
    const int SC = 3 * nj, RH = SC + 1, RL = SC + 4, S0 = SC + 7;
    for (int idx = tid; idx < nH + Lg; idx += nt) {
        double val = 0.0;
        if (idx < nH) {
            int ra, rb;
            if (!tri_decode(idx, Lg, ra, rb)) continue;
            if (ra < 3) {
                const int a = ra;
                if (rb < 3) {
                    val = (a == rb) ? gm(Gs, n, SC, SC) : 0.0;
                } else if (rb < 3 + 3 * nj) {   // (sc I)^T (-[u_k]x): eps_abc sum sc u_k,c
                    const int k = (rb - 3) / 3, b = (rb - 3) - 3 * k;
                    if (a != b) {
                        const double w = gm(Gs, n, SC, 3 * k + (3 - a - b));
                        val = (b == (a + 1) % 3) ? w : -w;
                    }
                } else {
                    val = gm(Gs, n, SC, S0 + a * K + (rb - 3 - 3 * nj));
                }
            } else if (ra < 3 + 3 * nj) {
                const int j = (ra - 3) / 3, a = (ra - 3) - 3 * j;
                if (rb < 3 + 3 * nj) {          // [u_j]x^T [u_k]x = (u_j . u_k) I - u_k u_j^T
                    const int k = (rb - 3) / 3, b = (rb - 3) - 3 * k;
                    val = -gm(Gs, n, 3 * k + a, 3 * j + b);
                    if (a == b)
                        val += gm(Gs, n, 3 * j, 3 * k) + gm(Gs, n, 3 * j + 1, 3 * k + 1) + gm(Gs, n, 3 * j + 2, 3 * k + 2);
                } else {                        // u_j x (sc S_m)
                    const int m = rb - 3 - 3 * nj, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                    val = gm(Gs, n, 3 * j + a1, S0 + a2 * K + m) - gm(Gs, n, 3 * j + a2, S0 + a1 * K + m);
                }
            } else {
                const int m = ra - 3 - 3 * nj, m2 = rb - 3 - 3 * nj;
                val = gm(Gs, n, S0 + m, S0 + m2) + gm(Gs, n, S0 + K + m, S0 + K + m2) +
                      gm(Gs, n, S0 + 2 * K + m, S0 + 2 * K + m2);
            }
        } else {
            const int ra = idx - nH;
            if (ra < 3) {
                val = gm(Gs, n, SC, RH + ra) + gm(Gs, n, SC, RL + ra);
            } else if (ra < 3 + 3 * nj) {       // u_j x rho
                const int j = (ra - 3) / 3, a = (ra - 3) - 3 * j, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                val = (gm(Gs, n, 3 * j + a1, RH + a2) + gm(Gs, n, 3 * j + a1, RL + a2)) -
                      (gm(Gs, n, 3 * j + a2, RH + a1) + gm(Gs, n, 3 * j + a2, RL + a1));
            } else {
                const int m = ra - 3 - 3 * nj;
                for (int c = 0; c < 3; ++c) val += gm(Gs, n, S0 + c * K + m, RH + c) + gm(Gs, n, S0 + c * K + m, RL + c);
            }
        }
        part[idx] = val;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32_lm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Gram matrix of a field-major fp32 tile T[nfp][ldt] in shared memory (cnt4 vertices, a multiple of 4) on the fp64 tensor
// path, expanded into the chunk's partial.  The tile memory is reused for the Gram matrix.
__device__ void gram_contract(const LmBuf& a, float* T, int ldt, int cnt4, int n, int nj, int K, double* part, bool wait_async,
                              unsigned long long& tp) {
    const int tid = threadIdx.x;
    // warp roles: 2x2-block tiles of the upper block triangle, round robin over the warps
    const int wid = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int n2 = (n + 1) >> 1, nwt = num_pairs(n2);
    int offa[kGramSlots][2], offb[kGramSlots][2], mask[kGramSlots], bis[kGramSlots], bjs[kGramSlots];
    double acc[kGramSlots][4][2];
#pragma unroll
    for (int s = 0; s < kGramSlots; ++s) {
        const int wt = wid + (kGramThreads / 32) * s;
        int ti = 0, tj = 0;
        mask[s] = 0;
        if (wt < nwt) {
            pair_to_blocks(wt, n2, ti, tj);
            const bool i1 = 2 * ti + 1 < n, j1 = 2 * tj + 1 < n;
            mask[s] = 1 | (j1 ? 2 : 0) | ((i1 && ti != tj) ? 4 : 0) | ((i1 && j1) ? 8 : 0);
        }
        bis[s] = 2 * ti;
        bjs[s] = 2 * tj;
        offa[s][0] = (8 * (2 * ti) + gid) * ldt + tig;
        offa[s][1] = (8 * min(2 * ti + 1, n - 1) + gid) * ldt + tig;
        offb[s][0] = (8 * (2 * tj) + gid) * ldt + tig;
        offb[s][1] = (8 * min(2 * tj + 1, n - 1) + gid) * ldt + tig;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[s][q][0] = acc[s][q][1] = 0.0;
    }
    if (wait_async) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    phase_lap(a.q, 14, tp);

#pragma unroll 2
    for (int k0 = 0; k0 < cnt4; k0 += 4) {
#pragma unroll
        for (int s = 0; s < kGramSlots; ++s) {
            if (mask[s]) {   // warp-uniform
                const double a0 = (double)T[offa[s][0] + k0], a1 = (double)T[offa[s][1] + k0];
                const double b0 = (double)T[offb[s][0] + k0], b1 = (double)T[offb[s][1] + k0];
                dmma884(acc[s][0][0], acc[s][0][1], a0, b0);
                if (mask[s] & 2) dmma884(acc[s][1][0], acc[s][1][1], a0, b1);
                if (mask[s] & 4) dmma884(acc[s][2][0], acc[s][2][1], a1, b0);
                if (mask[s] & 8) dmma884(acc[s][3][0], acc[s][3][1], a1, b1);
            }
        }
    }
    __syncthreads();   // every warp is done with the tile: reuse it for the Gram matrix
    phase_lap(a.q, 15, tp);
    double* Gs = reinterpret_cast<double*>(T);
#pragma unroll
    for (int s = 0; s < kGramSlots; ++s) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (mask[s] & (1 << q)) {
                const int bi = bis[s] + (q >> 1), bj = bjs[s] + (q & 1);
                *reinterpret_cast<double2*>(Gs + pair_index(bi, bj, n) * 64 + gid * 8 + 2 * tig) =
                    make_double2(acc[s][q][0], acc[s][q][1]);
            }
        }
    }
    __syncthreads();
    emit_partial(Gs, n, nj, K, part, tid, kGramThreads);
}

__device__ void gram_body(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, int c, unsigned char* smem_raw) {
    const int tid = threadIdx.x;
    const int K = M.K;
    unsigned long long tp = phase_begin(a.q);
    const int4 ch = a.chunks[(size_t)f * a.maxc + c];
    const int g = ch.x, start = ch.y, count = ch.z;
    const int nj = Pt.gnj[g];
    const int nf = rec_floats(nj, K), nfp = (nf + 7) & ~7, n = nfp >> 3;
    const int ldt = a.chunk_verts + 4;              // ldt % 32 == 4: conflict-free fragment loads
    const int cnt4 = (count + 3) & ~3, ng = cnt4 >> 2;
    float* T = reinterpret_cast<float*>(smem_raw);  // [nfp][ldt] field-major tile
    const float* recs = a.rec + (size_t)f * a.rec_stride * a.rec_rs + start;
    for (int e = tid; e < nf * ng; e += kGramThreads) {   // 16-byte granules; bytes past `count` are zero-filled
        const int q = e / ng, i = e - q * ng;
        const int valid = min(4, count - 4 * i) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32_lm(T + (size_t)q * ldt + 4 * i)),
                     "l"(recs + (size_t)q * a.rec_rs + 4 * i), "r"(valid) : "memory");
    }
    for (int e = tid; e < (nfp - nf) * cnt4; e += kGramThreads) T[(size_t)(nf + e / cnt4) * ldt + e % cnt4] = 0.f;
    asm volatile("cp.async.commit_group;" ::: "memory");

    gram_contract(a, T, ldt, cnt4, n, nj, K, a.part + ((size_t)f * a.maxc + c) * a.pstride, true, tp);
}

// Fused record + Gram task of the fp64 path of lm_flow_kernel: the records of one chunk (<= chunk_verts <= 256 matched vertices
// of one column group, one thread each) go straight into the shared-memory tile the contraction reads -- no d_rec round trip
// through HBM, no second queue hop.  Same arithmetic as rows_body + gram_body (record_vertex / gram_contract are shared).
__host__ __device__ inline size_t fused64_head_bytes(int tabD, int K) { return (((size_t)(tabD + ((K + 1) & ~1) + 32) * 8) + 127) & ~(size_t)127; }
__device__ void fused64_body(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, int c, bool cost_only, unsigned char* smem_raw) {
    const int tid = threadIdx.x;
    const int J = M.J, K = M.K;
    unsigned long long tp = phase_begin(a.q);
    const int4 ch = a.chunks[(size_t)f * a.maxc + c];
    const int g = ch.x, start = ch.y, count = ch.z;
    const int nj = Pt.gnj[g];
    const int nf = rec_floats(nj, K), nfp = (nf + 7) & ~7, n = nfp >> 3;
    const int ldt = a.chunk_verts + 4;
    const int cnt4 = (count + 3) & ~3;
    double* tab = reinterpret_cast<double*>(smem_raw);
    double* w = tab + a.tabD;
    double* scr = w + ((K + 1) & ~1);
    float* T = reinterpret_cast<float*>(smem_raw + fused64_head_bytes(a.tabD, K));
    const double* gtab = a.tab + (size_t)f * a.tabD;
    if ((a.tabD & 1) == 0) {   // 16-byte loads, all of a thread's loads in flight (as in rows_body)
        const double2* g2 = reinterpret_cast<const double2*>(gtab);
        double2* t2 = reinterpret_cast<double2*>(tab);
#pragma unroll 4
        for (int q = tid; q < (a.tabD >> 1); q += 256) t2[q] = __ldcg(g2 + q);
    } else {
        for (int q = tid; q < a.tabD; q += 256) tab[q] = ldg2(gtab + q);
    }
    for (int q = tid; q < K; q += 256) w[q] = ldg2(a.xt + (size_t)f * M.nx + 3 + 4 * J + q);
    const int v = (tid < count) ? (int)a.mlist[(size_t)f * a.rec_rs + start + tid] : (int)kNoVertex;
    if (!cost_only) {   // rows nf..nfp-1 and the columns of the last granule past `count` are zero (as the cp.async zero fill was)
        for (int e = tid; e < (nfp - nf) * cnt4; e += 256) T[(size_t)(nf + e / cnt4) * ldt + e % cnt4] = 0.f;
        if (tid < cnt4 && v == (int)kNoVertex)
            for (int q = 0; q < nf; ++q) T[(size_t)q * ldt + tid] = 0.f;
    }
    __syncthreads();
    phase_lap(a.q, 12, tp);
    double costv = 0.0;
    if (v != (int)kNoVertex)
        costv = record_vertex(M, Pt, a, f, v, g, T + tid, (size_t)ldt, cost_only, tab, w);
    const double cs = block_sum(costv, scr);
    if (tid == 0) a.cpart[(size_t)f * a.maxrb + c] = cs;
    phase_lap(a.q, 13, tp);
    if (cost_only) return;
    gram_contract(a, T, ldt, cnt4, n, nj, K, a.part + ((size_t)f * a.maxc + c) * a.pstride, false, tp);
}

__global__ void __launch_bounds__(kGramThreads, 2)
lm_gram_kernel(DevModel M, DevParts Pt, LmBuf a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int c = blockIdx.x, f = blockIdx.y;
    if (a.state[f].done || a.state[f].last || c >= a.state[f].nchunks) return;
    gram_body(M, Pt, a, f, c, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// fused record + Gram task of lm_flow_kernel<true> (the default path): J^T J on the 5th-generation tensor cores
// ---------------------------------------------------------------------------------------------
// One task = one chunk (<= 256 matched vertices of one column group), processed as sub-tiles of 128 vertices, two
// threads per vertex.
//
//  * Geometry in fp64: position x_v = sum_k w_k (G_k v0 + tau_k) (AvatarOptimizer.cpp:507-514), residual sum
//    rho' = c x - sum d (:632-639), cost.
//  * J^T J: the record fields [ u_j = 2 sc y_j | sc | sc S ] (:529-580 in closed form, see rows_body) are only ever
//    consumed as bf16 terms, so they are formed in fp32 from fp64-accurate differences and written STRAIGHT into
//    shared memory as the K-major SWIZZLE_128B operand tile of tcgen05.mma -- no Jacobian record goes to HBM.  Every field
//    is split into two bf16 terms (x = t0 + t1); the Gram matrix sum_v rec_v rec_v^T is the sum of the term products
//    t0 t0 + t0 t1 + t1 t0, accumulated by tcgen05.mma.kind::f16 (M = 128, N = fields rounded to 16, K = 16) into ONE
//    fp32 TMEM accumulator per CTA (128 columns, allocated once per persistent CTA).  The epilogue (tcgen05.ld) turns
//    the Gram matrix into the chunk's partial of J^T J exactly as the fp64 path does.
//    Accuracy (measured, tools/diag_h.py): the tensor cores' fp32 adds TRUNCATE, so J^T J comes out ~1.4e-6 low and up
//    to 5e-5 of the diagonal scale in near-degenerate twist directions.  Ten LM iterations amplify an H error of 1e-7
//    into ~3e-5 of parameter difference (they stop far from convergence), so this path is the fast OPTION
//    (AVB_JTJ_BF16_TENSOR: fits within ~3e-4 of the fp64 path, same objective to 1e-5), not the parity path.  An exact
//    integer variant (int8 slices, int32 TMEM accumulators, tools/ubench/umma_i8_test.cu: bit exact) was built and
//    measured too: three slices fit the TMEM budget of two CTAs per SM but their 22-bit fixed point per row bound is
//    coarser than bf16's floating split for the many fields far below their bound (H error 6e-4); parity-grade needs
//    five slices = five accumulators = the whole TMEM of an SM for one CTA.
//  * J^T r never touches the tensor cores and needs no Jacobian: by the chain rule through the skinning sum,
//        g_p = sum_v rho',   g_j = 2 sum_{k in subtree(j)} (M_k - (pos_j - o) x R_k),   g_shape,m = T_m + sum_k C_k[:,m] . R_k
//    with per-ASSIGNED-joint moments  R_k = sum w_k rho',  M_k = sum w_k (x^(k) - o) x rho'  and  T_m = sum Delta_v,m . (B_v^T rho'),
//    i.e. O(4) work per vertex instead of O(joints).  R and M are accumulated in 2^-36 fixed point with 64-bit integer
//    shared-memory atomics (order independent => bit reproducible), P and T with a fixed shuffle tree; the solve adds the
//    chunks in order and applies the subtree sums.  The gradient therefore has full fp64 accuracy -- the fixed point of
//    the iteration does not move; the tensor-core J^T J only steers the step (a third bf16 term costs twice the MMAs and
//    buys nothing: fp32 accumulation is the floor, tools/ubench/umma_gram_test.cu).
constexpr int kTcTerms = 2;                         // bf16 terms per fp32 field
constexpr int kTcSub = 128;                         // vertices per sub-tile: two 64-element swizzle atoms along K
constexpr int kTcRows = 80;                         // record fields of the tile (3 nj + 1 + 3 K <= 80)
constexpr int kTcAtomBytes = kTcRows * 128;         // one K atom: kTcRows rows of 128 B (64 bf16)
constexpr int kTcTermBytes = 2 * kTcAtomBytes;
constexpr int kTcTileBytes = kTcTerms * kTcTermBytes;
constexpr int kTcCols = 128;                        // TMEM columns per CTA (power of two >= kTcRows)
constexpr int kTcGw = 3 + kMaxK;                    // shuffle-reduced gradient sums per warp: P (3) | T (K)

__host__ __device__ inline int tc_fields(int nj, int K) { return 3 * nj + 1 + 3 * K; }
// gradient accumulators of a chunk partial: P (3) | R (J x 3) | M (J x 3) | T (K)
__host__ __device__ inline int tc_gacc(int J, int K) { return 3 + 6 * J + K; }

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    // start >> 4 | LBO (unused for swizzled K-major: 1) | SBO = 1024 B between 8-row groups | version 1 | SWIZZLE_128B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Field m of this thread's vertex into the term tiles.  Element (m, k) of a term tile lives at
//   (k >> 6) * atom + m * 128 + ((((k & 63) >> 3) ^ m) & 7) * 16 + (k & 7) * 2
// (8-row groups of 1024 B, rows of 128 B whose 16-byte chunks are XOR-swizzled with row & 7: what a SWIZZLE_128B
// shared-memory descriptor addresses); tile_kbase = shared address of the tile + (k >> 6) * atom + (k & 7) * 2 and
// kc = (k & 63) >> 3 are per thread.  32-bit shared addresses, st.shared: no generic address arithmetic.
__device__ __forceinline__ void tc_store(uint32_t tile_kbase, uint32_t kc, int m, float x) {
    const uint32_t addr = tile_kbase + ((uint32_t)m << 7) + (((kc ^ (uint32_t)m) & 7u) << 4);
#pragma unroll
    for (int t = 0; t < kTcTerms; ++t) {
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        if (t + 1 < kTcTerms) x -= __bfloat162float(b);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr + (uint32_t)(t * kTcTermBytes)), "h"(*reinterpret_cast<const unsigned short*>(&b)) : "memory");
    }
}
// sum over the 16 lanes of a half warp (the 16 vertices this half of the warp owns); every lane gets the sum
__device__ __forceinline__ double half_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ void tc_wait(uint64_t* mbar, uint32_t& phase) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32_lm(mbar)), "r"(phase) : "memory");
    phase ^= 1u;
}

// order-independent accumulation in 2^-36 fixed point.  Generic-address reduction on purpose: for a .shared address the
// compiler expands a 64-bit integer add into a compare-and-swap spin loop (ATOMS.CAST.SPIN.64), which collapses when the
// lanes of a warp hit the same joint; the generic form is a single RED.E.ADD.64.
__device__ __forceinline__ void fix_add(unsigned long long* p, double v) {
    asm volatile("red.add.u64 [%0], %1;" ::"l"(p), "l"((unsigned long long)__double2ll_rn(v * kFixScale)) : "memory");
}

struct TcSmem {
    unsigned char* tile;          // kTcTileBytes, 1024-byte aligned; later the fp64 Gram matrix
    double *G, *tau, *o, *w;      // fp64: global joint rotations, skinning translations, origin (root position), shape weights
    double *gsm, *scr;            // [8 warps][kTcGw] shuffle-reduced sums; block_sum scratch
    unsigned long long* acc;      // [J][6] fixed-point R_k | M_k
    float *Gf, *posr, *Cf;        // fp32 copies for the record fields: G, pos_j - o, C
    int* gjs;                     // [kMaxJ] the group's joints
};
__host__ __device__ inline size_t tc_smem_bytes(int J, int K) {
    // the tile must be followed by >= 6 KB of addressable shared memory: M = 128 makes the last K atom read 48 rows past
    // its 80 (their accumulator rows are ignored); the tables behind it provide that
    const size_t tail = (size_t)(12 * J + 4 + ((K + 1) & ~1) + 8 * kTcGw + 32 + 6 * J) * 8 + (size_t)(12 * J + 3 * J * K) * 4 + kMaxJ * 4;
    return 1024 + (size_t)kTcTileBytes + (tail < 6400 ? 6400 : tail) + 64;
}
__device__ inline TcSmem carve_tc(unsigned char* raw, int J, int K) {
    TcSmem S;
    // 1024-byte alignment (SWIZZLE_128B) by pointer arithmetic on the shared array -- no pointer -> integer -> pointer round
    // trip, so that the compiler keeps every access below in the shared address space (LDS / STS, 32-bit addresses)
    S.tile = raw + ((1024u - (smem_u32_lm(raw) & 1023u)) & 1023u);
    double* d = reinterpret_cast<double*>(S.tile + kTcTileBytes);
    S.G = d; d += 9 * J;
    S.tau = d; d += 3 * J;
    S.o = d; d += 4;
    S.w = d; d += (K + 1) & ~1;
    S.gsm = d; d += 8 * kTcGw;
    S.scr = d; d += 32;
    S.acc = reinterpret_cast<unsigned long long*>(d); d += 6 * J;
    float* fl = reinterpret_cast<float*>(d);
    S.Gf = fl; fl += 9 * J;
    S.posr = fl; fl += 3 * J;
    S.Cf = fl; fl += 3 * J * K;
    S.gjs = reinterpret_cast<int*>(fl);
    return S;
}

// chunk partial [ upper triangle of J^T J in the group's column order | P | R | M | T ] from the Gram matrix of
// [ u_j | sc | sc S ] and the gradient accumulators
__device__ void emit_partial_tc(const double* Gs, int n, int nj, int K, double* part, int tid, int nt) {
    const int Lg = 3 + 3 * nj + K, nH = tri_count(Lg);
    const int SC = 3 * nj, S0 = SC + 1;
    for (int idx = tid; idx < nH; idx += nt) {
        double val = 0.0;
        int ra, rb;
        if (!tri_decode(idx, Lg, ra, rb)) continue;
        if (ra < 3) {
            const int a = ra;
            if (rb < 3) {
                val = (a == rb) ? gm(Gs, n, SC, SC) : 0.0;
            } else if (rb < 3 + 3 * nj) {   // (sc I)^T (-[u_k]x): eps_abc sum sc u_k,c
                const int k = (rb - 3) / 3, b = (rb - 3) - 3 * k;
                if (a != b) {
                    const double w = gm(Gs, n, SC, 3 * k + (3 - a - b));
                    val = (b == (a + 1) % 3) ? w : -w;
                }
            } else {
                val = gm(Gs, n, SC, S0 + a * K + (rb - 3 - 3 * nj));
            }
        } else if (ra < 3 + 3 * nj) {
            const int j = (ra - 3) / 3, a = (ra - 3) - 3 * j;
            if (rb < 3 + 3 * nj) {          // [u_j]x^T [u_k]x = (u_j . u_k) I - u_k u_j^T
                const int k = (rb - 3) / 3, b = (rb - 3) - 3 * k;
                val = -gm(Gs, n, 3 * k + a, 3 * j + b);
                if (a == b)
                    val += gm(Gs, n, 3 * j, 3 * k) + gm(Gs, n, 3 * j + 1, 3 * k + 1) + gm(Gs, n, 3 * j + 2, 3 * k + 2);
            } else {                        // u_j x (sc S_m)
                const int m = rb - 3 - 3 * nj, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                val = gm(Gs, n, 3 * j + a1, S0 + a2 * K + m) - gm(Gs, n, 3 * j + a2, S0 + a1 * K + m);
            }
        } else {
            const int m = ra - 3 - 3 * nj, m2 = rb - 3 - 3 * nj;
            val = gm(Gs, n, S0 + m, S0 + m2) + gm(Gs, n, S0 + K + m, S0 + K + m2) +
                  gm(Gs, n, S0 + 2 * K + m, S0 + 2 * K + m2);
        }
        part[idx] = val;
    }
}

// tmem_d: the CTA's accumulator; mbar: its commit barrier; phase: parity the next commit completes (per thread copy)
// KT: number of shape keys when known at compile time (10 = SMPL: the shapedirs row of the vertex lives in registers and
// the shape loops unroll), 0 = any K at run time
template <int KT>
__device__ void fused_body(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, int c, bool cost_only,
                           unsigned char* smem_raw, uint32_t tmem_d, uint64_t* mbar, uint32_t& phase) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int J = M.J, K = KT > 0 ? KT : M.K;
    unsigned long long tp = phase_begin(a.q);
    const TcSmem S = carve_tc(smem_raw, J, K);
    const int4 ch = a.chunks[(size_t)f * a.maxc + c];
    const int g = ch.x, start = ch.y, count = ch.z;
    const int nj = Pt.gnj[g];
    const int nf = tc_fields(nj, K), Lg = group_L(nj, K);
    {
        const double* gtab = a.tab + (size_t)f * a.tabD;   // G | pos | tau | C of the trial point
        const double o0 = ldg2(gtab + 9 * J), o1 = ldg2(gtab + 9 * J + 1), o2 = ldg2(gtab + 9 * J + 2);
        for (int q = tid; q < 9 * J; q += 256) {
            const double v = ldg2(gtab + q);
            S.G[q] = v;
            S.Gf[q] = (float)v;
        }
        for (int q = tid; q < 3 * J; q += 256) {
            S.tau[q] = ldg2(gtab + 12 * J + q);
            const int cc = q % 3;
            S.posr[q] = (float)(ldg2(gtab + 9 * J + q) - (cc == 0 ? o0 : (cc == 1 ? o1 : o2)));
        }
        for (int q = tid; q < 3 * J * K; q += 256) S.Cf[q] = (float)ldg2(gtab + 15 * J + q);
        if (tid < 3) S.o[tid] = tid == 0 ? o0 : (tid == 1 ? o1 : o2);
        for (int q = tid; q < K; q += 256) S.w[q] = ldg2(a.xt + (size_t)f * M.nx + 3 + 4 * J + q);
        for (int q = tid; q < 8 * kTcGw; q += 256) S.gsm[q] = 0.0;
        for (int q = tid; q < 6 * J; q += 256) S.acc[q] = 0ull;
        for (int q = tid; q < nj; q += 256) S.gjs[q] = Pt.gjoints[g * kMaxJ + q];
    }
    __syncthreads();
    phase_lap(a.q, 12, tp);
    const int kk = (tid & 15) + 16 * wid;      // vertex slot inside the sub-tile; the two half warps split its fields
    const int h = (tid >> 4) & 1;
    const bool lead = (lane & 15) == 0;
    const uint32_t tkb = smem_u32_lm(S.tile) + (uint32_t)((kk >> 6) * kTcAtomBytes + (kk & 7) * 2), kc = (uint32_t)((kk & 63) >> 3);
    double* gw = S.gsm + wid * kTcGw;          // this warp's sums: one writer per entry (P: lane 0; T_m: the half that owns m)
    double costv = 0.0;
    const int nsub = (count + kTcSub - 1) / kTcSub;
    const int SC = 3 * nj, S0 = SC + 1;
    const double ox = S.o[0], oy = S.o[1], oz = S.o[2];
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
        const int li = sub * kTcSub + kk;
        int v = (li < count) ? (int)a.mlist[(size_t)f * a.rec_rs + start + li] : (int)kNoVertex;
        const bool valid = v != (int)kNoVertex;
        if (!valid) v = 0;
        // ---- fp64: shaped vertex, per-joint positions, position, residual sum, cost ----
        const float* sd = M.sd + (size_t)v * 3 * K;
        float sdr[KT > 0 ? 3 * KT : 2];   // KT > 0: the vertex's shapedirs row (3 x K floats, 8-byte aligned when 3 K is even)
        if (KT > 0) {
#pragma unroll
            for (int i = 0; i < (3 * KT) / 2; ++i) {
                const float2 t2 = __ldg(reinterpret_cast<const float2*>(sd) + i);
                sdr[2 * i] = t2.x;
                sdr[2 * i + 1] = t2.y;
            }
        }
#define AVB_SD(i) (KT > 0 ? sdr[(i)] : sd[(i)])
        double v0[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            double s0 = 0, s1 = 0;
            if (KT > 0) {
#pragma unroll
                for (int k = 0; k < KT; k += 2) {
                    s0 = fma((double)sdr[cc * KT + k], S.w[k], s0);
                    if (k + 1 < KT) s1 = fma((double)sdr[cc * KT + k + 1], S.w[k + 1], s1);
                }
            } else {
                int k = 0;
                for (; k + 1 < K; k += 2) {
                    s0 = fma((double)sd[cc * K + k], S.w[k], s0);
                    s1 = fma((double)sd[cc * K + k + 1], S.w[k + 1], s1);
                }
                if (k < K) s0 = fma((double)sd[cc * K + k], S.w[k], s0);
            }
            v0[cc] = M.vt[3 * (size_t)v + cc] + (s0 + s1);
        }
        const int n = valid ? (int)M.sk_n[v] : 0;
        double xk[AVB_MAX_ASSIGN_][3], wk[AVB_MAX_ASSIGN_];
        int jk[AVB_MAX_ASSIGN_];
        uint32_t mk[AVB_MAX_ASSIGN_];
        double x[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
            if (q < n) {
                const int k = M.sk_j[4 * (size_t)v + q];
                const double wt = M.sk_w[4 * (size_t)v + q];
                const double* Gk = S.G + 9 * k;
                jk[q] = k;
                wk[q] = wt;
                mk[q] = M.anc_mask[k];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    xk[q][cc] = Gk[3 * cc] * v0[0] + Gk[3 * cc + 1] * v0[1] + Gk[3 * cc + 2] * v0[2] + S.tau[3 * k + cc];
                    x[cc] += wt * xk[q][cc];
                }
            } else {
                jk[q] = 0; wk[q] = 0; mk[q] = 0;
                xk[q][0] = xk[q][1] = xk[q][2] = 0;
            }
        }
        const double cn = valid ? (double)a.cnt[(size_t)f * M.V + v] : 0.0;
        double rp[3] = {0, 0, 0};   // rho' = c x - sum d  (= sc * rho of rows_body)
        if (valid) {
            const unsigned long long* sumv = a.sum + 3 * ((size_t)f * M.V + v);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const double sr = (double)(long long)sumv[cc] * kFixInv;
                rp[cc] = cn * x[cc] - sr;
                if (h == 0) costv += x[cc] * (rp[cc] - sr);   // sum_i |x - d_i|^2 - sum_i |d_i|^2 = x . (c x - 2 s)
            }
        }
        if (cost_only) continue;
        // ---- fp64 gradient accumulators: this half's assigned joints q = h, h + 2 ----
        double t0 = 0, t1 = 0, t2 = 0;   // B^T rho' = sum_q w_q G_q^T rho'
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int q = 2 * qq + h;    // h is warp-half uniform: compile-time unrolled over qq, selected by h below
            const double wq = (q == 0) ? wk[0] : (q == 1 ? wk[1] : (q == 2 ? wk[2] : wk[3]));
            const int kq = (q == 0) ? jk[0] : (q == 1 ? jk[1] : (q == 2 ? jk[2] : jk[3]));
            const double xq0 = (q == 0) ? xk[0][0] : (q == 1 ? xk[1][0] : (q == 2 ? xk[2][0] : xk[3][0]));
            const double xq1 = (q == 0) ? xk[0][1] : (q == 1 ? xk[1][1] : (q == 2 ? xk[2][1] : xk[3][1]));
            const double xq2 = (q == 0) ? xk[0][2] : (q == 1 ? xk[1][2] : (q == 2 ? xk[2][2] : xk[3][2]));
            if (q < n) {
                const double r0 = wq * rp[0], r1 = wq * rp[1], r2 = wq * rp[2];
                const double a0 = xq0 - ox, a1 = xq1 - oy, a2 = xq2 - oz;
                unsigned long long* ac = S.acc + 6 * kq;
                fix_add(ac + 0, r0);
                fix_add(ac + 1, r1);
                fix_add(ac + 2, r2);
                fix_add(ac + 3, a1 * r2 - a2 * r1);
                fix_add(ac + 4, a2 * r0 - a0 * r2);
                fix_add(ac + 5, a0 * r1 - a1 * r0);
                const double* Gk = S.G + 9 * kq;
                t0 += Gk[0] * r0 + Gk[3] * r1 + Gk[6] * r2;
                t1 += Gk[1] * r0 + Gk[4] * r1 + Gk[7] * r2;
                t2 += Gk[2] * r0 + Gk[5] * r1 + Gk[8] * r2;
            }
        }
        t0 += __shfl_xor_sync(0xffffffffu, t0, 16);   // the other half's joints
        t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
        t2 += __shfl_xor_sync(0xffffffffu, t2, 16);
        {
            const double p0 = half_sum(rp[0]), p1 = half_sum(rp[1]), p2 = half_sum(rp[2]);
            if (lane == 0) { gw[0] += p0; gw[1] += p1; gw[2] += p2; }
        }
#pragma unroll
        for (int i = 0; 2 * i < (KT > 0 ? KT : K); ++i) {
            const int m = 2 * i + h;
            double dm0, dm1, dm2;
            if (KT > 0) {   // register row: select the half's entry (no dynamic register indexing)
                const int e1 = 2 * i + 1 < KT ? 2 * i + 1 : 2 * i;
                dm0 = h ? sdr[e1] : sdr[2 * i];
                dm1 = h ? sdr[KT + e1] : sdr[KT + 2 * i];
                dm2 = h ? sdr[2 * KT + e1] : sdr[2 * KT + 2 * i];
            } else {
                const int mm = m < K ? m : 0;
                dm0 = sd[mm]; dm1 = sd[K + mm]; dm2 = sd[2 * K + mm];
            }
            const double tm = half_sum(dm0 * t0 + dm1 * t1 + dm2 * t2);
            if (lead && m < K) gw[3 + m] += tm;
        }
        // ---- fp32 record fields -> bf16 term tiles ----
        float rq[AVB_MAX_ASSIGN_][3], wf[AVB_MAX_ASSIGN_], Bf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
            wf[q] = (float)wk[q];
            rq[q][0] = (float)(xk[q][0] - x[0]);
            rq[q][1] = (float)(xk[q][1] - x[1]);
            rq[q][2] = (float)(xk[q][2] - x[2]);
            const float* Gq = S.Gf + 9 * jk[q];
#pragma unroll
            for (int e = 0; e < 9; ++e) Bf[e] = fmaf(wf[q], Gq[e], Bf[e]);   // wf == 0 for q >= n
        }
        const float xo0 = (float)(x[0] - ox), xo1 = (float)(x[1] - oy), xo2 = (float)(x[2] - oz);
        const float scf = (float)sqrt(cn), s2f = 2.f * scf;
        if (sub > 0) tc_wait(mbar, phase);   // the previous sub-tile's MMAs have read the tile: it may be rewritten
        if (h == 0) tc_store(tkb, kc, SC, scf);
        // rotation fields u_j = 2 sc (sum_{q in desc(j)} w_q (x^(q) - x) + W_j (x - pos_j))
#pragma unroll 1
        for (int i = 0; 2 * i < nj; ++i) {
            const int gi = 2 * i + h;
            if (gi < nj) {
                const int j = S.gjs[gi];
                float y0 = 0, y1 = 0, y2 = 0, W = 0;
#pragma unroll
                for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
                    const float wq = ((mk[q] >> j) & 1u) ? wf[q] : 0.f;
                    W += wq;
                    y0 = fmaf(wq, rq[q][0], y0);
                    y1 = fmaf(wq, rq[q][1], y1);
                    y2 = fmaf(wq, rq[q][2], y2);
                }
                tc_store(tkb, kc, 3 * gi, s2f * fmaf(W, xo0 - S.posr[3 * j], y0));
                tc_store(tkb, kc, 3 * gi + 1, s2f * fmaf(W, xo1 - S.posr[3 * j + 1], y1));
                tc_store(tkb, kc, 3 * gi + 2, s2f * fmaf(W, xo2 - S.posr[3 * j + 2], y2));
            }
        }
        // shape fields sc (B Delta_v + sum_q w_q C_q) (AvatarOptimizer.cpp:568-580)
#pragma unroll
        for (int i = 0; 2 * i < (KT > 0 ? KT : K); ++i) {
            const int m = 2 * i + h;
            if (m < K) {
                float d0, d1, d2;
                if (KT > 0) {
                    const int e1 = 2 * i + 1 < KT ? 2 * i + 1 : 2 * i;
                    d0 = h ? sdr[e1] : sdr[2 * i];
                    d1 = h ? sdr[KT + e1] : sdr[KT + 2 * i];
                    d2 = h ? sdr[2 * KT + e1] : sdr[2 * KT + 2 * i];
                } else {
                    d0 = sd[m]; d1 = sd[K + m]; d2 = sd[2 * K + m];
                }
                float e0 = Bf[0] * d0 + Bf[1] * d1 + Bf[2] * d2;
                float e1 = Bf[3] * d0 + Bf[4] * d1 + Bf[5] * d2;
                float e2 = Bf[6] * d0 + Bf[7] * d1 + Bf[8] * d2;
#pragma unroll
                for (int q = 0; q < AVB_MAX_ASSIGN_; ++q) {
                    const float* Cq = S.Cf + jk[q] * 3 * K;   // wf == 0 for q >= n (jk == 0: a valid table row)
                    e0 = fmaf(wf[q], Cq[m], e0);
                    e1 = fmaf(wf[q], Cq[K + m], e1);
                    e2 = fmaf(wf[q], Cq[2 * K + m], e2);
                }
                tc_store(tkb, kc, S0 + m, scf * e0);
                tc_store(tkb, kc, S0 + K + m, scf * e1);
                tc_store(tkb, kc, S0 + 2 * K + m, scf * e2);
            }
        }
        // ---- tensor cores: D += sum over term pairs of T_ta T_tb^T over this sub-tile's vertices ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> tensor-core reads
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int N = (nf + 15) & ~15;
            // instruction descriptor: D = F32, A = B = BF16, K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const int ksteps = (min(kTcSub, count - sub * kTcSub) + 15) >> 4;   // unused slots hold zeros (sc = 0) up to a multiple of 16
            const uint32_t base = smem_u32_lm(S.tile);
            uint32_t accum = sub > 0 ? 1u : 0u;
#pragma unroll 1
            for (int ta = 0; ta < kTcTerms; ++ta)
#pragma unroll 1
                for (int tb = 0; ta + tb < kTcTerms; ++tb)
#pragma unroll 1
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t koff = (uint32_t)((ks >> 2) * kTcAtomBytes + (ks & 3) * 32);
                        const uint64_t da = tc_desc(base + ta * kTcTermBytes + koff), db = tc_desc(base + tb * kTcTermBytes + koff);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                        accum = 1u;
                    }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32_lm(mbar)) : "memory");
        }
    }
    phase_lap(a.q, 13, tp);
    const double cs = block_sum(costv, S.scr);
    if (tid == 0) a.cpart[(size_t)f * a.maxrb + c] = cs;
    if (cost_only) return;
    tc_wait(mbar, phase);                                  // every MMA of the chunk has retired
    asm volatile("tcgen05.fence::after_thread_sync;");
    phase_lap(a.q, 14, tp);
    // ---- epilogue: TMEM (lane = field m, column = field) -> fp64 Gram matrix (upper block triangle) over the tile ----
    const int n = (nf + 7) >> 3;
    double* Gs = reinterpret_cast<double*>(S.tile);
    if ((wid & 3) < 3) {
        const int m = 32 * (wid & 3) + lane, bi = m >> 3;
        const int cb0 = (wid >> 2) ? (n >> 1) : 0, cb1 = (wid >> 2) ? n : (n >> 1);
#pragma unroll 1
        for (int cb = cb0; cb < cb1; ++cb) {
            uint32_t r[8];
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)(8 * cb);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (bi < n && cb >= bi) {
                double* p = Gs + pair_index(bi, cb, n) * 64 + (m & 7) * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) p[q] = (double)__uint_as_float(r[q]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    double* part = a.part + ((size_t)f * a.maxc + c) * a.pstride;
    {   // gradient accumulators behind the J^T J triangle: P | R | M | T
        double* ga = part + tri_count(Lg);
        const int GA = tc_gacc(J, K);
        for (int i = tid; i < GA; i += 256) {
            double val;
            if (i < 3 || i >= 3 + 6 * J) {
                const int e = i < 3 ? i : 3 + (i - 3 - 6 * J);
                val = 0;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) val += S.gsm[w8 * kTcGw + e];   // warp partials in warp order
            } else {
                const int e = i - 3, k = e < 3 * J ? e / 3 : (e - 3 * J) / 3, cc = e % 3;
                val = (double)(long long)S.acc[6 * k + (e < 3 * J ? cc : 3 + cc)] * kFixInv;
            }
            ga[i] = val;
        }
    }
    __syncthreads();
    emit_partial_tc(Gs, n, nj, K, part, tid, 256);
    phase_lap(a.q, 15, tp);
}

// ---------------------------------------------------------------------------------------------
// lm_solve_kernel
// ---------------------------------------------------------------------------------------------
constexpr int kNBsq = 64;   // 8 x 8 diagonal block of the Cholesky panels
struct SolveSmem {
    double *xs, *xt, *tb, *Hs, *gs, *glo, *gcur, *delta, *dd, *wscr, *aa, *ycomp, *dcomp, *scr;
    int* iscr;
};
// table scratch of the solve: the joint tables of the retraction, earlier the joint rotations of the basis change and
// the column-major panel of the factorisation (8 x ((P + 4) & ~3) + 8 doubles)
// packed lower triangle, row major: entry (i, k), k <= i, at tri(i) + k.  The solve keeps J^T J (+ priors, + damping) and
// its Cholesky factor in this form: 29 KB instead of 58 KB for P = 85, which is what lets three or four CTAs share an SM.
__host__ __device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
// leading dimension of the column-major Cholesky panel Lp[c][row - t0]: rows padded to whole 8-row tiles, = 4 (mod 8) so
// that the fragment loads of the tensor-path trailing update (lane -> column lane & 3, row lane >> 2) are bank-conflict free
__host__ __device__ __forceinline__ int chol_ldp(int R) { return ((R + 7) & ~7) + 4; }
__host__ __device__ inline int solve_tb_doubles(int J, int K, int C) {
    // early phases: joint rotations G (9 J) | gradient moments (3 + 6 J + K) | pos (3 J) | C (3 J K);
    // factorisation: column-major panel (8 x chol_ldp(P + 1) + 8).  The joint tables of the retraction alias the matrix.
    // (The prior's dcomp | ycomp have their own scratch: the prior is evaluated next to the partial reduction.)
    const int P = 3 + 3 * J + K, D = 3 * (J - 1);
    int t = 9 * J + ((tc_gacc(J, K) + 1) & ~1) + 3 * J + 3 * J * K;
    int pnl = 8 * chol_ldp(P + 1) + 8;
    const int ninv = kNBsq * ((P + 7) >> 3);   // block inverses of the back substitution
    pnl = pnl > ninv ? pnl : ninv;
    (void)C; (void)D;
    return ((t > pnl ? t : pnl) + 1) & ~1;
}
__host__ __device__ inline int solve_h_doubles(int J, int K) {
    const int P = 3 + 3 * J + K, h = tri(P + 1) + 2, t = tables_doubles(J, K, true);
    return ((h > t ? h : t) + 1) & ~1;
}
__host__ __device__ inline size_t solve_smem_bytes(int J, int K, int C) {
    const int P = 3 + 3 * J + K, nx = 3 + 4 * J + K, D = 3 * (J - 1);
    size_t d = 2 * ((nx + 1) & ~1) + solve_tb_doubles(J, K, C) + solve_h_doubles(J, K) + 5 * ((P + 1) & ~1) + 8 * kNBsq + ((D + 1) & ~1) + 64 +
               2 * (size_t)(C > 0 ? C : 1) * ((D + 8) & ~7);
    return d * 8 + (8 + 4 * kMaxGroups + 8) * 4 + 128;   // + iscr: 8 scalars | the group runs staged by the solve ([kMaxGroups][4])
}
__device__ inline SolveSmem carve_solve(unsigned char* raw, const DevModel& M) {
    SolveSmem S;
    const int P = M.P, nx = M.nx, D = 3 * (M.J - 1);
    double* d = reinterpret_cast<double*>(raw);
    S.xs = d; d += (nx + 1) & ~1;
    S.xt = d; d += (nx + 1) & ~1;
    S.tb = d; d += solve_tb_doubles(M.J, M.K, M.gmmC);
    S.Hs = d; d += solve_h_doubles(M.J, M.K);   // packed lower triangle + the augmented right-hand-side row P
    S.gs = d; d += (P + 1) & ~1;
    S.glo = d; d += (P + 1) & ~1;
    S.gcur = d; d += (P + 1) & ~1;
    S.delta = d; d += (P + 1) & ~1;
    S.dd = d; d += (P + 1) & ~1;
    S.wscr = d; d += 8 * kNBsq;
    S.aa = d; d += (D + 1) & ~1;
    S.scr = d; d += 64;
    S.dcomp = d; d += (size_t)(M.gmmC > 0 ? M.gmmC : 1) * ((D + 8) & ~7);
    S.ycomp = d; d += (size_t)(M.gmmC > 0 ? M.gmmC : 1) * ((D + 8) & ~7);
    S.iscr = reinterpret_cast<int*>(d);
    return S;
}

// CTA-wide blocked right-looking Cholesky of the leading P x P lower triangle of W (PACKED lower, row major: tri(i) + k),
// carried through R >= P rows: rows P..R-1 enter as right-hand sides b^T and leave as (L^-1 b)^T, i.e. the forward
// substitution comes for free.  dinv receives 1 / L_jj.  Returns false (uniformly) when a pivot is not positive.
//
// The factorisation is a chain of P dependent pivots; everything else is throughput.  So warp 0 runs the chain AHEAD of
// the trailing update (measured, tools/ubench/chol_ubench.cu: 54.0 k -> 31.5 k cycles for P = 85 on one CTA):
//   warp 0, panel j:  its eight rows of the panel (x = w L11^-T) -> bar.arrive -> update of the NEXT diagonal block from
//                     those rows (registers) -> its factorisation (one row per lane, shuffles) -> L11 of panel j + 1
//   warps 1..7:       the other rows of the panel -> bar.sync -> trailing update S -= X X^T on 8 x 8 tiles with the fp64
//                     tensor path (mma.m8n8k4, two k-steps per tile, two tiles in flight), every tile except warp 0's
//   one CTA barrier per panel.  Pivot step: a_c -= (a_k a_ck) / d with 1 / d from the fp64 MUFU seed (2^-20) and one
//   cubic Newton step (three dependent fp64 operations); the square roots are taken after the chain.
constexpr int kNB = 8;
// 1 / d for a pivot in the guarded range: 2^-20 seed, one cubic step (error e^3 = 2^-60)
__device__ __forceinline__ double pivot_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}
// 1 / sqrt(d): 2^-19 seed, one cubic step and one linear fix-up (full double accuracy up to rounding)
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d * y, y, 1.0);
    y = fma(y, fma(0.375 * e, e, 0.5 * e), y);
    e = fma(-d * y, y, 1.0);
    return fma(0.5 * y, e, y);
}
#ifdef AVB_UBENCH_PHASES
__device__ long long g_chol_cyc[4];
#define CHOL_T(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long t_ = clock64(); g_chol_cyc[k] += t_ - tph; tph = t_; } } while (0)
#else
#define CHOL_T(k) do { } while (0)
#endif
// one warp, lane & 7 = row: factor the nb x nb block whose row r (entries c <= r; rows >= nb: identity) the lane holds in a[].
// Writes L to W, the row-scaled copy Lw[c][k] = L_ck / L_cc (diagonal: 1 / L_cc) and dinv.
__device__ __forceinline__ bool diag_factor(double (&a)[kNB], double* W, int j0, int nb, double* Lw, double* dinv, int lane) {
    const int r = lane & 7;
    bool ok = true;
    double mydiag = 1.0;
#pragma unroll
    for (int k = 0; k < kNB; ++k) {
        const double d = __shfl_sync(0xffffffffu, a[k], k);
        if (!(d > 1e-300) || !(d < 1e300)) ok = false;   // not positive, not finite, or outside the seed's range
        const double inv = pivot_rcp(d);
        if (r == k) mydiag = d;
#pragma unroll
        for (int c = k + 1; c < kNB; ++c) {
            const double ack = __shfl_sync(0xffffffffu, a[k], c);
            a[c] = fma(-(a[k] * ack), inv, a[c]);
        }
    }
    if (!ok) return false;   // uniform over the warp
    const double myrs = pivot_rsqrt(mydiag);
#pragma unroll
    for (int k = 0; k < kNB; ++k) {
        const double rsk = __shfl_sync(0xffffffffu, myrs, k);
        a[k] *= rsk;   // L_rk = a_rk / sqrt(d_k); k == r: sqrt(d_r)
    }
    __syncwarp();   // the duplicate lanes (8..31) have read the rows that lanes 0..7 overwrite now
    if (lane < kNB) {
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            Lw[r * kNB + k] = (k < r) ? a[k] * myrs : ((k == r) ? myrs : 0.0);
            if (r < nb && k <= r) W[tri(j0 + r) + j0 + k] = a[k];
        }
    }
    if (lane < nb) dinv[j0 + lane] = myrs;
    return true;
}
// row i of the panel: x L11^T = W[i][j0 .. j0 + nb), column-oriented (one fma on the critical path per column)
__device__ __forceinline__ void panel_row(double* W, int i, int j0, int nb, int t0, const double* Lw, double* Lp, int ldp,
                                          double (&x)[kNB], bool store, bool sync_warp) {
    double* wi = W + tri(i) + j0;
#pragma unroll
    for (int c = 0; c < kNB; ++c) x[c] = (c < nb) ? wi[c] * Lw[c * kNB + c] : 0.0;
#pragma unroll
    for (int k = 0; k < kNB - 1; ++k)
#pragma unroll
        for (int c = k + 1; c < kNB; ++c) x[c] = fma(-x[k], Lw[c * kNB + k], x[c]);
    if (sync_warp) __syncwarp();   // warp 0: duplicate lanes read the row another lane is about to overwrite
    if (store) {
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
            if (c < nb) wi[c] = x[c];
            Lp[c * ldp + (i - t0)] = x[c];   // columns beyond nb hold zeros
        }
    }
}
// wscr: Lw[2][64] | flag[2] (needs 2 * 64 + 1 doubles); panel: 8 * chol_ldp(R) + 4 doubles
__device__ bool aug_cholesky(double* W, int P, int R, double* dinv, double* wscr, double* panel) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
#ifdef AVB_UBENCH_PHASES
    long long tph = clock64();
#endif
    int* flag = reinterpret_cast<int*>(wscr + 2 * kNB * kNB);
    double* Lp = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(panel) + 31) & ~(uintptr_t)31);
    const int ldp = chol_ldp(R);
    const int r = lane & 7;
    for (int e = tid; e < kNB * ldp; e += nthr) Lp[e] = 0.0;
    if (wid == 0) {
        const int nb0 = min(kNB, P);
        double a[kNB];
#pragma unroll
        for (int c = 0; c < kNB; ++c) a[c] = (r < nb0 && c <= r) ? W[tri(r) + c] : ((c == r) ? 1.0 : 0.0);
        const bool ok = diag_factor(a, W, 0, nb0, wscr, dinv, lane);
        if (lane == 0) flag[0] = ok ? 1 : 0;
    }
    __syncthreads();
    CHOL_T(0);
    int pb = 0;
    for (int j0 = 0; j0 < P; j0 += kNB, pb ^= 1) {
        if (!flag[pb]) return false;   // written before the barrier every thread has just passed: uniform
        const int nb = min(kNB, P - j0), t0 = j0 + nb;
        const double* Lw = wscr + pb * (kNB * kNB);
        const int mrow = R - t0, mcol = P - t0;
        if (wid == 0) {
            double x[kNB];
            const int i = t0 + r;
            const bool have = i < R;
            panel_row(W, min(i, R - 1), j0, nb, t0, Lw, Lp, ldp, x, have && lane < kNB, true);
            if (!have) {
#pragma unroll
                for (int c = 0; c < kNB; ++c) x[c] = 0.0;
            }
            __syncwarp();
            asm volatile("bar.arrive 1, %0;" ::"r"(nthr) : "memory");   // the panel rows of this warp are in Lp
            if (mcol > 0) {
                // rows t0 .. t0 + 7 x columns t0 .. t0 + nbn - 1: the next diagonal block and (at the end of the matrix) the
                // right-hand-side rows that share its 8-row tile
                const int nbn = min(kNB, mcol);
                double a[kNB];
#pragma unroll
                for (int c = 0; c < kNB; ++c) a[c] = (have && c < nbn && c <= r) ? W[tri(i) + t0 + c] : 0.0;
#pragma unroll
                for (int k = 0; k < kNB; ++k) {
                    const double4 lo = *reinterpret_cast<const double4*>(Lp + k * ldp), hi = *reinterpret_cast<const double4*>(Lp + k * ldp + 4);
                    const double xc[kNB] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int c = 0; c < kNB; ++c) a[c] = fma(-x[k], xc[c], a[c]);
                }
                __syncwarp();     // (duplicate lanes read the same rows of W)
                if (r >= nbn) {   // not a row of the diagonal block: store, and present an identity row to the factorisation
#pragma unroll
                    for (int c = 0; c < kNB; ++c) {
                        if (have && lane < kNB && c < nbn) W[tri(i) + t0 + c] = a[c];
                        a[c] = (c == r) ? 1.0 : 0.0;
                    }
                }
                const bool ok = diag_factor(a, W, t0, nbn, wscr + (pb ^ 1) * (kNB * kNB), dinv, lane);
                if (lane == 0) flag[pb ^ 1] = ok ? 1 : 0;
            } else if (lane == 0) {
                flag[pb ^ 1] = 1;
            }
        } else {
            for (int i = t0 + kNB + (tid - 32); i < R; i += nthr - 32) {
                double x[kNB];
                panel_row(W, i, j0, nb, t0, Lw, Lp, ldp, x, true, false);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
            if (mcol > 0) {
                // S[i][k] -= sum_c X[i][c] X[k][c] on 8 x 8 tiles of the lower triangle, row-block order (row block ti holds
                // min(ti + 1, Tc) tiles); tile 0 is warp 0's.  The walk is incremental: no division, no square root.
                const int Tb = (mrow + 7) >> 3, Tc = (mcol + 7) >> 3;
                const int ntri = Tc * (Tc + 1) / 2, ntile = ntri + (Tb - Tc) * Tc;
                const int g = lane >> 2, q = lane & 3, step = nwarp - 1;
                auto advance = [&](int& ti, int& tk, int n) {
                    tk += n;
                    for (int len = min(ti + 1, Tc); tk >= len; len = min(ti + 1, Tc)) {
                        tk -= len;
                        ++ti;
                    }
                };
                int ti0 = 0, tk0 = 0;
                advance(ti0, tk0, wid);
                for (int t = wid; t < ntile; t += 2 * step) {
                    int ti1 = ti0, tk1 = tk0;
                    const bool two = t + step < ntile;
                    if (two) advance(ti1, tk1, step);
                    const double a00 = -Lp[q * ldp + 8 * ti0 + g], a01 = -Lp[(4 + q) * ldp + 8 * ti0 + g];
                    const double b00 = Lp[q * ldp + 8 * tk0 + g], b01 = Lp[(4 + q) * ldp + 8 * tk0 + g];
                    const double a10 = -Lp[q * ldp + 8 * ti1 + g], a11 = -Lp[(4 + q) * ldp + 8 * ti1 + g];
                    const double b10 = Lp[q * ldp + 8 * tk1 + g], b11 = Lp[(4 + q) * ldp + 8 * tk1 + g];
                    const int i0 = t0 + 8 * ti0 + g, k0 = t0 + 8 * tk0 + 2 * q;
                    const int i1 = t0 + 8 * ti1 + g, k1 = t0 + 8 * tk1 + 2 * q;
                    const bool v00 = i0 < R && k0 < P && k0 <= i0, v01 = i0 < R && k0 + 1 < P && k0 + 1 <= i0;
                    const bool v10 = two && i1 < R && k1 < P && k1 <= i1, v11 = two && i1 < R && k1 + 1 < P && k1 + 1 <= i1;
                    double* w0 = W + tri(min(i0, R - 1)) + k0;
                    double* w1 = W + tri(min(i1, R - 1)) + k1;
                    double c00 = v00 ? w0[0] : 0.0, c01 = v01 ? w0[1] : 0.0, c10 = v10 ? w1[0] : 0.0, c11 = v11 ? w1[1] : 0.0;
                    double e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;
                    dmma884(c00, c01, a00, b00);
                    dmma884(e00, e01, a01, b01);
                    dmma884(c10, c11, a10, b10);
                    dmma884(e10, e11, a11, b11);
                    if (v00) w0[0] = c00 + e00;
                    if (v01) w0[1] = c01 + e01;
                    if (v10) w1[0] = c10 + e10;
                    if (v11) w1[1] = c11 + e11;
                    ti0 = ti1;
                    tk0 = tk1;
                    if (t + 2 * step < ntile) advance(ti0, tk0, step);
                }
            }
        }
        __syncthreads();
        CHOL_T(1);
    }
    return flag[pb] != 0;
}

// x = L^-T y, blocked by the 8 x 8 diagonal blocks of the factor (85 dependent steps become 11):
//   block_inverses   every warp inverts diagonal blocks (lane c: column c of N = L_bb^-1 by forward substitution, 1 / L_rr from
//                    dinv) and stores them transposed, NT[b][c][r] = N[r][c]   (needs 64 * ceil(P / 8) doubles; CTA-wide, the
//                    caller synchronises)
//   warp_back_solve  one warp, blocks last to first: x_b = N_b^T y_b (eight independent dot products, no chain), then
//                    y_e -= sum_c L[j0 + c][e] x_{j0 + c} for the entries e < j0 (rows of L are contiguous in e).
// In place in shared memory (x and y may alias).  Measured against the column-by-column loop: 10.5 -> 6.4 us per solve (a variant with the
// update unrolled over four predicated entries per lane measured 11.4: a single warp pays for every instruction it issues).
__device__ void block_inverses(const double* L, const double* dinv, int P, double* NT) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int c = lane & 7;
#pragma unroll 1
    for (int b = wid; b * kNB < P; b += nwarp) {
        const int j0 = b * kNB, nb = min(kNB, P - j0);
        double n[kNB];   // n[r] = N[r][c], r >= c
#pragma unroll
        for (int r = 0; r < kNB; ++r) {
            double acc = 0.0;
            const double* Lr = L + tri(j0 + min(r, nb - 1)) + j0;
#pragma unroll
            for (int m = 0; m < kNB; ++m)
                if (m < r) acc = fma(Lr[m], (m >= c) ? n[m] : 0.0, acc);
            const double di = dinv[j0 + min(r, nb - 1)];
            n[r] = (r == c) ? di : ((r > c && r < nb) ? -acc * di : 0.0);
        }
        if (lane < kNB) {
#pragma unroll
            for (int r = 0; r < kNB; ++r) NT[b * (kNB * kNB) + c * kNB + r] = n[r];
        }
    }
}
__device__ void warp_back_solve(const double* L, const double* NT, int P, const double* y, double* x) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int e = lane; e < P; e += 32) x[e] = y[e];
    __syncwarp();
#pragma unroll 1
    for (int b = (P - 1) / kNB; b >= 0; --b) {
        const int j0 = b * kNB, nb = min(kNB, P - j0);
        const int c = lane & 7;
        const double* nt = NT + b * (kNB * kNB) + c * kNB;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < kNB; r += 2) {   // N[r][c] = 0 for r < c and for the rows beyond nb
            s0 = fma(nt[r], (r < nb) ? x[j0 + r] : 0.0, s0);
            s1 = fma(nt[r + 1], (r + 1 < nb) ? x[j0 + r + 1] : 0.0, s1);
        }
        const double xc = s0 + s1;
        __syncwarp();
        if (lane < nb) x[j0 + lane] = xc;
        double xb[kNB];
#pragma unroll
        for (int q = 0; q < kNB; ++q) xb[q] = __shfl_sync(0xffffffffu, xc, q);
#pragma unroll 1
        for (int e = lane; e < j0; e += 32) {
            double acc = x[e];
#pragma unroll
            for (int q = 0; q < kNB; ++q)
                if (q < nb) acc = fma(-L[tri(j0 + q) + e], xb[q], acc);
            x[e] = acc;
        }
        __syncwarp();
    }
}

// Pose prior at the trial point xt (AvatarOptimizer.cpp:661-692, GaussianMixture.cpp:95-114), by threads t = 0..nt-1 (whole
// warps, named barrier 3): the component of least p_c = 1/2 (x - mu_c)^T Sigma_c^-1 (x - mu_c) - consts_log[c] (first minimum
// wins), its cost 1/2 sbp^2 p_c -> *cost_out and y_c = Sigma_c^-1 (x - mu_c) of EVERY component -> ycomp (row stride (D + 8) & ~7).
// The caller synchronises before it reads the results.
__device__ void prior_eval(const DevModel& M, const double* xt, double sbp, int t, int nt, double* aa, double* dcomp, double* ycomp,
                           int* best_out, double* cost_out) {
    const int J = M.J;
    // pose prior (AvatarOptimizer.cpp:661-692, GaussianMixture.cpp:95-114): component of least
    // p_c = 1/2 (x - mu_c)^T Sigma_c^-1 (x - mu_c) - consts_log[c], its value and y = Sigma^-1 (x - mu)
    const int D = M.gmmD, C = M.gmmC, Dp = (D + 8) & ~7;   // row stride: zero padded to 8, slot D of ycomp holds p_c
#pragma unroll 1
    for (int j = 1 + t; j < J; j += nt) {  // Eigen AngleAxisd(Quaterniond)
        const double* q = xt + 3 + 4 * j;
        double nn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        double s = 0.0;
        if (nn != 0.0) {
            const double ang = 2.0 * atan2(nn, fabs(q[3]));
            if (q[3] < 0) nn = -nn;
            s = ang / nn;
        }
        aa[3 * (j - 1)] = q[0] * s;
        aa[3 * (j - 1) + 1] = q[1] * s;
        aa[3 * (j - 1) + 2] = q[2] * s;
    }
    asm volatile("bar.sync 3, %0;" ::"r"(nt) : "memory");
#pragma unroll 1
    for (int i = t; i < C * Dp; i += nt) {
        const int cc = i / Dp, k = i - cc * Dp;
        dcomp[i] = (k < D) ? aa[k] - M.gmm_mean[cc * D + k] : 0.0;
    }
    asm volatile("bar.sync 3, %0;" ::"r"(nt) : "memory");
#pragma unroll 1
    for (int i = t; i < C * D; i += 2 * nt) {  // y_c = Sigma_c^-1 (x - mu_c), two rows per thread at a time
        // the precision matrices are symmetric (symmetrised at load): walk column r so that consecutive
        // threads read consecutive addresses; sixteen loads in flight per thread (the walk is L2-latency bound)
        const int i2 = i + nt;
        const bool two = i2 < C * D;
        const int ca = i / D, ra = i - ca * D;
        const int cb = two ? i2 / D : ca, rb = two ? i2 - cb * D : ra;
        const double* Pa = M.gmm_prec + (size_t)ca * D * D + ra;
        const double* Pb = M.gmm_prec + (size_t)cb * D * D + rb;
        const double* da = dcomp + ca * Dp;
        const double* db = dcomp + cb * Dp;
        double sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0;
#pragma unroll 1
        for (int k = 0; k < D; k += 8) {
            double ta[8], tb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ta[u] = (k + u < D) ? __ldg(Pa + (size_t)(k + u) * D) : 0.0;
                tb[u] = (k + u < D) ? __ldg(Pb + (size_t)(k + u) * D) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                sa0 = fma(ta[u], da[k + u], sa0);          // dcomp rows are zero padded to a multiple of 8
                sa1 = fma(ta[u + 1], da[k + u + 1], sa1);
                sb0 = fma(tb[u], db[k + u], sb0);
                sb1 = fma(tb[u + 1], db[k + u + 1], sb1);
            }
        }
        ycomp[ca * Dp + ra] = sa0 + sa1;
        if (two) ycomp[cb * Dp + rb] = sb0 + sb1;
    }
    asm volatile("bar.sync 3, %0;" ::"r"(nt) : "memory");
    // p_c, one warp per component; first minimum wins (strict <)
#pragma unroll 1
    for (int cc = t >> 5; cc < C; cc += nt >> 5) {
        double sp = 0;
#pragma unroll 1
        for (int k = t & 31; k < D; k += 32) sp += dcomp[cc * Dp + k] * ycomp[cc * Dp + k];
        sp = 0.5 * warp_sum(sp);
        if ((t & 31) == 0) ycomp[cc * Dp + D] = sp;
    }
    asm volatile("bar.sync 3, %0;" ::"r"(nt) : "memory");
    if (t == 0) {
        double bestp = 1.79769313486231570e308, bestsq = 0;
        int best = 0;
#pragma unroll 1
        for (int cc = 0; cc < C; ++cc) {
            const double sq = ycomp[cc * Dp + D];
            const double pc = sq - M.gmm_clog[cc];
            if (pc < bestp) {
                bestp = pc;
                bestsq = sq;
                best = cc;
            }
        }
        *best_out = best;
        *cost_out = 0.5 * sbp * sbp * (bestsq - M.gmm_clog[best]);
    }
}

// the prior as a task of lm_flow_kernel (small batches): evaluate at the frame's trial point, leave cost | component | y in HBM
__device__ void prior_body(const DevModel& M, const LmBuf& a, int f, unsigned char* smem_raw) {
    const int tid = threadIdx.x, D = M.gmmD, C = M.gmmC, Dp = (D + 8) & ~7, nx = M.nx;
    double* xt = reinterpret_cast<double*>(smem_raw);
    double* aa = xt + ((nx + 1) & ~1);
    double* dcomp = aa + ((D + 1) & ~1);
    double* ycomp = dcomp + (size_t)C * Dp;
    double* res = ycomp + (size_t)C * Dp;   // [0] cost, [1] component (as int)
    for (int i = tid; i < nx; i += 256) xt[i] = ldg2(a.xt + (size_t)f * nx + i);
    const double sbp = ldg2(&a.state[f].sbp);
    __syncthreads();
    prior_eval(M, xt, sbp, tid, 256, aa, dcomp, ycomp, reinterpret_cast<int*>(res + 1), res);
    __syncthreads();
    double* out = a.prior_out + (size_t)f * a.prior_stride;
    const int best = *reinterpret_cast<int*>(res + 1);
    if (tid == 0) {
        out[0] = res[0];
        out[1] = (double)best;
    }
    for (int r = tid; r < D; r += 256) out[2 + r] = ycomp[best * Dp + r];
}
__host__ __device__ inline size_t prior_task_smem(int nx, int C, int D) {
    return (size_t)(((nx + 1) & ~1) + ((D + 1) & ~1) + 2 * (size_t)(C > 0 ? C : 1) * ((D + 8) & ~7) + 4) * 8;
}

// returns true when the frame has finished (uniform over the CTA)
__device__ bool solve_body(const DevModel& M, const DevParts& Pt, const LmBuf& a, int f, unsigned char* smem_raw) {
    const int tid = threadIdx.x;
    LmState& gst = a.state[f];
    const int P = M.P, nx = M.nx, J = M.J, K = M.K;
    SolveSmem S = carve_solve(smem_raw, M);
    const int nTri = tri(P);   // packed entries of the P x P lower triangle
    double* Hcur = a.Hcur + (size_t)f * P * P;   // (packed: the first tri(P) doubles of the frame's slot)
    double* gcur_g = a.gcur + (size_t)f * P;
    const LmState st = load_state(&gst);  // snapshot (only thread 0 writes it back at the end)
    unsigned long long tp = phase_begin(a.q);

#pragma unroll 1
    for (int i = tid; i < nx; i += kSolveThreads) {
        S.xs[i] = ldg2(a.x + (size_t)f * nx + i);
        S.xt[i] = ldg2(a.xt + (size_t)f * nx + i);
    }
#pragma unroll 1
    for (int i = tid; i < P; i += kSolveThreads) {
        S.gcur[i] = ldg2(gcur_g + i);
        S.gs[i] = 0.0;
    }
#pragma unroll 4
    for (int i = tid; i < nTri; i += kSolveThreads) S.Hs[i] = 0.0;
    double* Gs9 = S.tb;   // global joint rotations of the trial point (the table scratch is free until the retraction)
#pragma unroll 1
    for (int i = tid; i < 9 * J; i += kSolveThreads) Gs9[i] = ldg2(a.tab + (size_t)f * a.tabD + i);
    // the frame's chunk runs and the static group tables: one round trip here instead of one per group in the reduction
    int* s_run = S.iscr + 8;                 // [numGroups][4]: first chunk, #chunks, entries of a partial, offset into gdest
    if (tid < Pt.numGroups) {
        const int2 run = a.gruns[(size_t)f * kMaxGroups + tid];
        const int Lg = group_L(Pt.gnj[tid], K);
        s_run[4 * tid] = run.x;
        s_run[4 * tid + 1] = run.y;
        s_run[4 * tid + 2] = tri_count(Lg) + (a.tensor ? 0 : Lg);   // tensor path: the gradient comes from the moment accumulators
        s_run[4 * tid + 3] = Pt.gdoff[tid];
    }
    __syncthreads();
    phase_lap(a.q, 4, tp);
    // ---- two independent latency-bound jobs run side by side on the two halves of the CTA (named barriers 2 and 3):
    //      threads 0..127 reduce the chunk partials, threads 128..255 evaluate the pose prior at the trial point ----
    const bool last = st.last != 0;   // cost-only evaluation: it decides the final accept / reject, no step follows
    const double sbp = st.sbp, sbs = st.sbs;
    const bool do_prior = sbp > 0.0 && M.gmmC > 0;
    const bool prior_ready = do_prior && a.prior_task;   // evaluated by a prior task of this evaluation (lm_flow_kernel, small batches)
    const int nred = (do_prior && !prior_ready) ? kSolveThreads / 2 : kSolveThreads;
    if (prior_ready) {
        const double* po = a.prior_out + (size_t)f * a.prior_stride;
        const int D = M.gmmD, Dp = (D + 8) & ~7, best = (int)ldg2(po + 1);
        if (tid == 0) {
            S.iscr[0] = best;
            S.scr[40] = ldg2(po);
        }
#pragma unroll 1
        for (int r = tid; r < D; r += kSolveThreads) S.ycomp[best * Dp + r] = ldg2(po + 2 + r);
    }
    if (tid < nred) {
        // reduce the chunk partials, group by group (chunks of a group share one layout), chunks in order
#pragma unroll 1
        for (int g = 0; g < (last ? 0 : Pt.numGroups); ++g) {
            const int2 run = make_int2(s_run[4 * g], s_run[4 * g + 1]);
            if (run.y <= 0) continue;   // uniform
            const int nE = s_run[4 * g + 2];
            const int* dest = Pt.gdest + s_run[4 * g + 3];  // static scatter table: entry -> packed tangent index (gradient: nTri + column)
            const double* part = a.part + ((size_t)f * a.maxc + run.x) * a.pstride;
            constexpr int kB = 8;   // entries per thread per pass (one pass covers a group of up to 9 joints on 128 threads)
#pragma unroll 1
            for (int i0 = tid; i0 < nE; i0 += nred * kB) {
                double val[kB];
                int dst[kB];
#pragma unroll
                for (int u = 0; u < kB; ++u) {
                    const int idx = i0 + u * nred;
                    val[u] = (idx < nE) ? ldg2(part + idx) : 0.0;
                    dst[u] = (idx < nE) ? __ldg(dest + idx) : -1;
                }
#pragma unroll 1
                for (int cc = 1; cc < run.y; ++cc) {   // further chunks of the group, in chunk order
#pragma unroll
                    for (int u = 0; u < kB; ++u) {
                        const int idx = i0 + u * nred;
                        if (idx < nE) val[u] += ldg2(part + (size_t)cc * a.pstride + idx);
                    }
                }
#pragma unroll
                for (int u = 0; u < kB; ++u) {
                    if (dst[u] >= 0) {
                        double* o = dst[u] < nTri ? S.Hs + dst[u] : S.gs + (dst[u] - nTri);
                        *o += val[u];
                    }
                }
            }
            asm volatile("bar.sync 2, %0;" ::"r"(nred) : "memory");
        }
        phase_lap(a.q, 5, tp);   // (thread 0 is on this side: 'reduce' is the reduction alone, the wait for the prior side counts as 'prior')
    } else {
        prior_eval(M, S.xt, sbp, tid - nred, kSolveThreads - nred, S.aa, S.dcomp, S.ycomp, &S.iscr[0], &S.scr[40]);
    }
    __syncthreads();
    phase_lap(a.q, 7, tp);
    double csum = 0.0;   // cost partials in record-block order, eight loads in flight
    {
        const int nb = (a.tensor || a.fused) ? st.nchunks : (st.nslots + 255) >> 8;   // one cost partial per fused task / per record block
#pragma unroll 1
        for (int b = 0; b < nb; b += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = (b + u < nb) ? ldg2(a.cpart + (size_t)f * a.maxrb + b + u) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (b + u < nb) csum += t[u];
        }
    }
    double cost_t = 0.5 * (csum + st.Qsum);
    __syncthreads();
    if (a.tensor && !last) {
        // ---- J^T r from the moment accumulators of the fused tasks (see fused_body), chunks added in chunk order:
        //      g_p = P,  g_j = 2 sum_{k in subtree(j)} (M_k - (pos_j - o) x R_k),  g_shape,m = T_m + sum_k C_k[:,m] . R_k ----
        const int GA = tc_gacc(J, K);
        double* ga = S.tb + 9 * J;          // behind the joint rotations G (table scratch)
        double* posj = ga + ((GA + 1) & ~1);
        double* Ctab = posj + 3 * J;
#pragma unroll 1
        for (int i = tid; i < GA; i += kSolveThreads) {
            double sacc = 0.0;
#pragma unroll 1
            for (int cc = 0; cc < st.nchunks; ++cc) {
                const int gc = a.chunks[(size_t)f * a.maxc + cc].x;
                sacc += ldg2(a.part + ((size_t)f * a.maxc + cc) * a.pstride + tri_count(group_L(Pt.gnj[gc], K)) + i);
            }
            ga[i] = sacc;
        }
#pragma unroll 1
        for (int i = tid; i < 3 * J; i += kSolveThreads) posj[i] = ldg2(a.tab + (size_t)f * a.tabD + 9 * J + i);
#pragma unroll 1
        for (int i = tid; i < 3 * J * K; i += kSolveThreads) Ctab[i] = ldg2(a.tab + (size_t)f * a.tabD + 15 * J + i);
        __syncthreads();
        const double* Rk = ga + 3;
        const double* Mk = ga + 3 + 3 * J;
#pragma unroll 1
        for (int i = tid; i < 3; i += kSolveThreads) S.gs[i] = ga[i];
#pragma unroll 1
        for (int i = tid; i < 3 * J; i += kSolveThreads) {
            const int j = i / 3, cc = i - 3 * j, c1 = (cc + 1) % 3, c2 = (cc + 2) % 3;
            const double pj1 = posj[3 * j + c1] - posj[c1], pj2 = posj[3 * j + c2] - posj[c2];   // o = root position
            double sacc = 0.0;
#pragma unroll 1
            for (int k = 0; k < J; ++k)
                if ((M.anc_mask[k] >> j) & 1u) sacc += Mk[3 * k + cc] - (pj1 * Rk[3 * k + c2] - pj2 * Rk[3 * k + c1]);
            S.gs[3 + i] = 2.0 * sacc;
        }
#pragma unroll 1
        for (int m = tid; m < K; m += kSolveThreads) {
            double sacc = ga[3 + 6 * J + m];
#pragma unroll 1
            for (int k = 0; k < J; ++k)
                sacc += Ctab[(3 * k) * K + m] * Rk[3 * k] + Ctab[(3 * k + 1) * K + m] * Rk[3 * k + 1] + Ctab[(3 * k + 2) * K + m] * Rk[3 * k + 2];
            S.gs[3 + 3 * J + m] = sacc;
        }
        __syncthreads();
    }
    phase_lap(a.q, 5, tp);
    // ---- eta -> delta coordinates: H = T^T Ht T, g = T^T gt, T_j = G_parent(j) at the trial point ----
    if (!last) {
        // One pass over the blocks of the lower triangle, every block owned by one thread (in place):
        //   joint x joint (3 x 3, A >= B; the diagonal blocks are symmetric and stored as six entries), joint rows x p columns
        //   (left factor only), shape rows x joint columns (1 x 3 strips, right factor only).  Root joint, p, shape: T = I.
        const int nJJ = (J * (J + 1)) >> 1, nItems = nJJ + J + K * J;
#pragma unroll 1
        for (int it = tid; it < nItems; it += kSolveThreads) {
            if (it < nJJ) {
                int A = (int)((sqrtf(8.f * (float)it + 1.f) - 1.f) * 0.5f);
                while (tri(A + 1) <= it) ++A;
                while (tri(A) > it) --A;
                const int B = it - tri(A);
                if (A == 0) continue;                       // root x root: identity on both sides
                const double* TA = Gs9 + 9 * M.parent[A];
                const int ra = 3 + 3 * A, cb = 3 + 3 * B;
                double X[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int rr = ra + r, cc = cb + c;
                        X[r][c] = S.Hs[rr >= cc ? tri(rr) + cc : tri(cc) + rr];   // mirrors the diagonal block
                    }
                double Y[3][3];   // T_A^T X
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Y[r][c] = TA[r] * X[0][c] + TA[3 + r] * X[1][c] + TA[6 + r] * X[2][c];
                if (B > 0) {      // ... T_B
                    const double* TB = Gs9 + 9 * M.parent[B];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const double y0 = Y[r][0], y1 = Y[r][1], y2 = Y[r][2];
                        Y[r][0] = y0 * TB[0] + y1 * TB[3] + y2 * TB[6];
                        Y[r][1] = y0 * TB[1] + y1 * TB[4] + y2 * TB[7];
                        Y[r][2] = y0 * TB[2] + y1 * TB[5] + y2 * TB[8];
                    }
                }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (A != B || c <= r) S.Hs[tri(ra + r) + cb + c] = Y[r][c];
            } else if (it < nJJ + J) {
                const int A = it - nJJ;
                if (A == 0) continue;
                const double* TA = Gs9 + 9 * M.parent[A];
                const int ra = 3 + 3 * A;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double x0 = S.Hs[tri(ra) + c], x1 = S.Hs[tri(ra + 1) + c], x2 = S.Hs[tri(ra + 2) + c];
                    S.Hs[tri(ra) + c] = TA[0] * x0 + TA[3] * x1 + TA[6] * x2;
                    S.Hs[tri(ra + 1) + c] = TA[1] * x0 + TA[4] * x1 + TA[7] * x2;
                    S.Hs[tri(ra + 2) + c] = TA[2] * x0 + TA[5] * x1 + TA[8] * x2;
                }
            } else {
                const int q = it - nJJ - J, m = q / J, B = q - m * J;
                if (B == 0) continue;
                const double* TB = Gs9 + 9 * M.parent[B];
                double* h = S.Hs + tri(3 + 3 * J + m) + 3 + 3 * B;
                const double h0 = h[0], h1 = h[1], h2 = h[2];
                h[0] = h0 * TB[0] + h1 * TB[3] + h2 * TB[6];
                h[1] = h0 * TB[1] + h1 * TB[4] + h2 * TB[7];
                h[2] = h0 * TB[2] + h1 * TB[5] + h2 * TB[8];
            }
        }
#pragma unroll 1
        for (int j = 1 + tid; j < J; j += kSolveThreads) {
            const double* Gp = Gs9 + 9 * M.parent[j];
            double* gg = S.gs + 3 + 3 * j;
            const double g0 = gg[0], g1 = gg[1], g2 = gg[2];
            gg[0] = Gp[0] * g0 + Gp[3] * g1 + Gp[6] * g2;
            gg[1] = Gp[1] * g0 + Gp[4] * g1 + Gp[7] * g2;
            gg[2] = Gp[2] * g0 + Gp[5] * g1 + Gp[8] * g2;
        }
    }
    __syncthreads();
    phase_lap(a.q, 6, tp);
    // ---- pose prior, second half: add the winning component (evaluated above, next to the partial reduction) ----
    const double* w = S.xt + 3 + 4 * J;
    if (do_prior) {
        const int D = M.gmmD, Dp = (D + 8) & ~7;
        const int best = S.iscr[0];
        const double hb = 0.5 * sbp * sbp;
        if (!last) {   // H += hb Sigma_best^-1, stored zero-padded to the P x P tangent layout: one flat pass, no index maths
            const double* Pf = M.gmm_pfull + (size_t)best * nTri;   // Sigma^-1 zero padded to the tangent layout, packed lower
#pragma unroll 1
            for (int i0 = tid; i0 < nTri; i0 += 8 * kSolveThreads) {
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = (i0 + u * kSolveThreads < nTri) ? __ldg(Pf + i0 + u * kSolveThreads) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (i0 + u * kSolveThreads < nTri) S.Hs[i0 + u * kSolveThreads] = fma(hb, t[u], S.Hs[i0 + u * kSolveThreads]);
            }
        }
#pragma unroll 1
        for (int r = tid; r < (last ? 0 : D); r += kSolveThreads) S.gs[6 + r] += hb * S.ycomp[best * Dp + r];
        __syncthreads();   // the flat pass above touched every entry of H, the shape prior below owns some of them
        cost_t += S.scr[40];
    }
    // ---- shape prior (AvatarOptimizer.cpp:708-723) ----
    if (sbs > 0.0) {
        double sq = 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) sq += w[k] * w[k];
        cost_t += 0.5 * sbs * sbs * sq;
#pragma unroll 1
        for (int k = tid; k < K; k += kSolveThreads) {
            S.Hs[tri(3 + 3 * J + k) + 3 + 3 * J + k] += sbs * sbs;
            S.gs[3 + 3 * J + k] += sbs * sbs * w[k];
        }
    }
    __syncthreads();
    phase_lap(a.q, 7, tp);
    if (a.dump_cost) {  // avb_debug_evaluate: report the objective at the evaluation point and stop
        if (tid == 0) {
            a.dump_cost[f] = cost_t;
            gst.done = 1;
        }
#pragma unroll 1
        for (int i = tid; i < P; i += kSolveThreads) a.dump_grad[(size_t)f * P + i] = S.gs[i];
#pragma unroll 1
        for (int i = tid; i < P * P; i += kSolveThreads) {   // debug dump: the full symmetric matrix
            const int r = i / P, c = i - r * P;
            a.dump_H[(size_t)f * P * P + i] = S.Hs[r >= c ? tri(r) + c : tri(c) + r];
        }
        return true;
    }

    // ---- Levenberg-Marquardt step control (Ceres-1.14-style trust region; DESIGN.md "solver") ----
    double cost = st.cost, radius = st.radius, decrease = st.decrease, initial_cost = st.initial_cost;
    int iters = st.iters, accepted = st.accepted;
    bool done = false, have_cur_in_smem = false, take_point = false;
    double gm = 0;   // max |g| at the evaluation point (every thread: a short rolled loop beats a reduction here)
#pragma unroll 1
    for (int i = 0; i < (last ? 0 : P); ++i) gm = fmax(gm, fabs(S.gs[i]));
    if (last) gm = 1.0;   // no gradient in a cost-only evaluation; the frame ends after it in any case
    if (st.evals == 0) {
        // first evaluation: the trial point is the start point
        cost = initial_cost = cost_t;
        take_point = true;
        done = !(gm > 1e-10) || !isfinite(cost);
    } else {
        // finish iteration `iters`: accept or reject the trial point
        const double rho = (cost - cost_t) / st.model_change;
        if (isfinite(cost_t) && rho > 1e-3) {
            ++accepted;
            double dn = 0, xn = 0;
#pragma unroll 1
            for (int i = 0; i < nx; ++i) {
                const double dd = S.xt[i] - S.xs[i];
                dn += dd * dd;
                xn += S.xs[i] * S.xs[i];
            }
            const double cost_change = cost - cost_t, cost_old = cost;
            take_point = true;
            cost = cost_t;
            const double t3 = 2.0 * rho - 1.0;
            radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - t3 * t3 * t3));
            decrease = 2.0;
            if (sqrt(dn) <= 1e-8 * (sqrt(xn) + 1e-8)) done = true;
            if (fabs(cost_change) <= a.function_tolerance * cost_old) done = true;
            if (!(gm > 1e-10)) done = true;
        } else {
            radius /= decrease;
            decrease *= 2.0;
            if (radius < 1e-32) done = true;
        }
    }
    if (take_point) {   // the evaluation point becomes the current point (uniform over the CTA)
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < nx; i += kSolveThreads) S.xs[i] = S.xt[i];
#pragma unroll 4
        for (int i = tid; i < (last ? 0 : nTri); i += kSolveThreads) Hcur[i] = S.Hs[i];
#pragma unroll 1
        for (int i = tid; i < (last ? 0 : P); i += kSolveThreads) S.gcur[i] = S.gs[i];
        __syncthreads();
        have_cur_in_smem = !last;
    }
    auto trace_now = [&]() {
        if (a.trace && iters >= 1) {
            const int slot = min(iters - 1, a.trace_cap - 1);
#pragma unroll 1
            for (int i = tid; i < nx; i += kSolveThreads) a.trace[((size_t)f * a.trace_cap + slot) * nx + i] = S.xs[i];
        }
    };
    if (st.evals > 0) trace_now();
    // ---- next iteration(s): damped solve until a usable step exists ----
    double model_change = 0.0;
    bool have_step = false;
    phase_lap(a.q, 8, tp);
    while (!done && iters < a.max_iters && !have_step) {
        ++iters;
        if (!have_cur_in_smem) {
#pragma unroll 1
            for (int i = tid; i < nTri; i += kSolveThreads) S.Hs[i] = ldg2(Hcur + i);
            __syncthreads();
        }
        have_cur_in_smem = false;  // the factorisation below overwrites S.Hs
        // W = Hcur + D,  D_jj = clamp(s^2 h_jj, 1e-6, 1e32) / (s^2 radius),  s = 1 / (1 + sqrt(h_jj)); row P = -g^T
#pragma unroll 1
        for (int j = tid; j < P; j += kSolveThreads) {
            const double h = S.Hs[tri(j) + j];
            const double sj = 1.0 / (1.0 + sqrt(h));
            const double dj = fmin(fmax(sj * sj * h, 1e-6), 1e32) / (sj * sj * radius);
            S.dd[j] = dj;
            S.Hs[tri(j) + j] = h + dj;
            S.Hs[nTri + j] = -S.gcur[j];
        }
        __syncthreads();
        bool ok = aug_cholesky(S.Hs, P, P + 1, S.glo, S.wscr, S.tb);
        phase_lap(a.q, 9, tp);
        if (ok) {
            block_inverses(S.Hs, S.glo, P, S.tb);   // (the panel of the factorisation is dead: its place takes the block inverses)
            __syncthreads();
            if (tid < 32) warp_back_solve(S.Hs, S.tb, P, S.Hs + nTri, S.delta);
            __syncthreads();
            // model_cost_change = -delta^T (g + 1/2 H delta) with (H + D) delta = -g  =>  1/2 delta^T (D delta - g)
            double part = 0;
#pragma unroll 1
            for (int r = tid; r < P; r += kSolveThreads) part += 0.5 * S.delta[r] * (S.dd[r] * S.delta[r] - S.gcur[r]);
            model_change = block_sum(part, S.scr);
            ok = model_change > 0.0 && isfinite(model_change);
        }
        if (ok) {
            have_step = true;
        } else {
            radius /= decrease;
            decrease *= 2.0;
            if (radius < 1e-32) done = true;
            trace_now();
        }
    }
    if (!have_step) done = true;
    phase_lap(a.q, 10, tp);
    if (!done) {
        // retraction: p, w additive; q <- dq (x) q, |delta| is the half angle (AvatarOptimizer.cpp:123-143)
#pragma unroll 1
        for (int i = tid; i < 3; i += kSolveThreads) S.xt[i] = S.xs[i] + S.delta[i];
#pragma unroll 1
        for (int k = tid; k < K; k += kSolveThreads) S.xt[3 + 4 * J + k] = S.xs[3 + 4 * J + k] + S.delta[3 + 3 * J + k];
#pragma unroll 1
        for (int j = tid; j < J; j += kSolveThreads) {
            const double* d = S.delta + 3 + 3 * j;
            const double* q = S.xs + 3 + 4 * j;
            double* o = S.xt + 3 + 4 * j;
            const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (nd > 0.0) {
                double sn, aw;
                sincos(nd, &sn, &aw);
                const double sdd = sn / nd;
                const double ax = sdd * d[0], ay = sdd * d[1], az = sdd * d[2];
                o[0] = aw * q[0] + ax * q[3] + ay * q[2] - az * q[1];
                o[1] = aw * q[1] + ay * q[3] + az * q[0] - ax * q[2];
                o[2] = aw * q[2] + az * q[3] + ax * q[1] - ay * q[0];
                o[3] = aw * q[3] - ax * q[0] - ay * q[1] - az * q[2];
            } else {
                o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
            }
        }
        __syncthreads();
        Tables T = carve_tables(S.Hs, J, K, true);   // the factor is dead: the joint tables of the trial point take its place
        build_tables(M, S.xt, T, true);
        double* tab = a.tab + (size_t)f * a.tabD;
#pragma unroll 1
        for (int i = tid; i < 9 * J; i += kSolveThreads) tab[i] = T.G[i];
#pragma unroll 1
        for (int i = tid; i < 3 * J; i += kSolveThreads) {
            tab[9 * J + i] = T.pos[i];
            tab[12 * J + i] = T.tau[i];
        }
#pragma unroll 1
        for (int i = tid; i < 3 * J * K; i += kSolveThreads) tab[15 * J + i] = T.C[i];
#pragma unroll 1
        for (int i = tid; i < nx; i += kSolveThreads) a.xt[(size_t)f * nx + i] = S.xt[i];
    }
#pragma unroll 1
    for (int i = tid; i < nx; i += kSolveThreads) a.x[(size_t)f * nx + i] = S.xs[i];
#pragma unroll 1
    for (int i = tid; i < P; i += kSolveThreads) gcur_g[i] = S.gcur[i];
    if (tid == 0) {
        gst.cost = cost;
        gst.radius = radius;
        gst.decrease = decrease;
        gst.initial_cost = initial_cost;
        gst.model_change = model_change;
        gst.iters = iters;
        gst.accepted = accepted;
        gst.evals = st.evals + 1;
        gst.last = (!done && iters >= a.max_iters) ? 1 : 0;   // the pending trial is the last one: cost only
        gst.done = done ? 1 : 0;
        FrameStats& fs = a.stats[f];
        // an iteration whose trial point is still to be evaluated is counted once that evaluation is reduced
        fs.iterations = done ? iters : iters - 1;
        fs.accepted_steps = accepted;
        fs.initial_cost = initial_cost;
        fs.final_cost = cost;
        if (!isfinite(cost)) fs.status = 4;
    }
    phase_lap(a.q, 11, tp);
    return done;
}

__global__ void __launch_bounds__(kSolveThreads, 2)
lm_solve_kernel(DevModel M, DevParts Pt, LmBuf a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (a.state[blockIdx.x].done) return;
    solve_body(M, Pt, a, blockIdx.x, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// lm_flow_kernel: the whole inner solve (every evaluation of every frame) as one persistent data-flow kernel
// ---------------------------------------------------------------------------------------------
// CTAs take tasks from a global queue: rows(f, block) -> gram(f, chunk) -> solve(f) -> rows(f, ...) of the next
// evaluation.  A task is pushed only when its inputs are complete (per-frame countdowns), so no CTA ever waits on
// another one; the CTA that finishes a frame's last chunk runs the frame's solve itself.  The latency-bound solves
// of some frames overlap the throughput-bound record and Gram tasks of others, and the 3 x (1 + maxItersPerICP)
// kernel boundaries of the staged schedule disappear.  Results are identical to the staged kernels: the same
// bodies run on the same data in the same per-frame order.
// TC = true: the default path, fused record + Gram tasks on tcgen05 (fused_body); TC = false: fp64 DMMA Gram from fp32
// records in HBM (rows_body + gram_body).
// OCC: CTAs per SM the kernel is compiled for (register budget 128 / 80 / 64 per thread); the host picks the variant.
template <bool TC, int OCC>
__global__ void __launch_bounds__(256, OCC)
lm_flow_kernel(DevModel M, DevParts Pt, LmBuf a) {
    extern __shared__ __align__(1024) unsigned char smem_flow[];
    unsigned char* const smem_raw = smem_flow;
    __shared__ int s_task, s_last;
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x;
    uint32_t tmem_d = 0, mma_phase = 0;
    if (TC) {   // the CTA's fp32 accumulator in tensor memory, held for the lifetime of the persistent CTA
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32_lm(&s_tmem)), "r"(kTcCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_lm(&s_mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        tmem_d = s_tmem;
    }
    unsigned long long t_prev = 0;
    if (a.q.prof && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_prev));
    auto lap = [&](int cls) {   // thread 0: CTA time per task class (avb_set_profiling)
        if (a.q.prof) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicAdd(a.q.prof + cls, t - t_prev);
            t_prev = t;
        }
    };
    for (;;) {
        if (tid == 0) {
            s_task = flow_pop(a.q);
            lap(3);
            if (s_task == -1) flow_leave(a.q);
        }
        __syncthreads();
        const int task = s_task;
        if (task == -1) break;
        const int type = (int)((unsigned)task >> 30), f = (task >> 12) & 0x3FFFF, idx = task & 0xFFF;
        if (type == kTaskPrior) {
            prior_body(M, a, f, smem_raw);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_last = atomicSub(&a.q.gram_left[f], 1) == 1;
                lap(2);
            }
        } else if (TC) {
            const bool cost_only = ldg2(&a.state[f].last) != 0;
            if (M.K == 10)
                fused_body<10>(M, Pt, a, f, idx, cost_only, smem_raw, tmem_d, &s_mbar, mma_phase);
            else
                fused_body<0>(M, Pt, a, f, idx, cost_only, smem_raw, tmem_d, &s_mbar, mma_phase);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_last = atomicSub(&a.q.gram_left[f], 1) == 1;
                lap(1);
            }
        } else if (type == kTaskFused) {
            const bool cost_only = ldg2(&a.state[f].last) != 0;
            fused64_body(M, Pt, a, f, idx, cost_only, smem_raw);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_last = atomicSub(&a.q.gram_left[f], 1) == 1;
                lap(1);
            }
        } else if (type == kTaskRows) {
            const int nslots = ldg2(&a.state[f].nslots);
            const bool cost_only = ldg2(&a.state[f].last) != 0;
            rows_body(M, Pt, a, f, idx, nslots, cost_only, smem_raw);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_last = 0;
                if (atomicSub(&a.q.rows_left[f], 1) == 1) {
                    __threadfence();
                    if (cost_only) {   // no Gram tasks: the records were the evaluation (gram_left counts them as one)
                        s_last = atomicSub(&a.q.gram_left[f], 1) == 1;
                    } else {           // (gram_left was set when the evaluation was opened, flow_push_eval)
                        flow_push(a.q, kTaskGram, f, ldg2(&a.state[f].nchunks));
                    }
                }
                lap(0);
            }
        } else {
            gram_body(M, Pt, a, f, idx, smem_raw);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_last = atomicSub(&a.q.gram_left[f], 1) == 1;
                lap(1);
            }
        }
        __syncthreads();
        if (s_last) {   // the frame's evaluation is complete: this CTA runs its solve
            __threadfence();
            const bool done = solve_body(M, Pt, a, f, smem_raw);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                if (done) {
                    atomicSub(&a.q.ctrl[2], 1u);
                } else {
                    flow_push_eval(a, f, ldg2(&a.state[f].nslots), ldg2(&a.state[f].nchunks), ldg2(&a.state[f].last) != 0,
                                   a.prior_task && ldg2(&a.state[f].sbp) > 0.0);
                }
                lap(2);
            }
        }
        __syncthreads();
    }
    if (TC) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTcCols));
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
size_t lm_prep_smem(const DevModel& M) {
    return (size_t)(((M.nx + 1) & ~1) + tables_doubles(M.J, M.K, true) + 64) * 8 + 64 * 4 + 64;
}
size_t lm_rows_smem(const DevModel& M) {
    return (size_t)(tab_doubles(M.J, M.K) + ((M.K + 1) & ~1) + 32) * 8 + (kMaxGroups + 2) * 4 + 256 * 4 + 64;
}
size_t lm_gram_smem(const DevModel& M, int max_nj, int chunk_verts, bool tensor) {
    const int nf = rec_floats(max_nj, M.K), nfp = (nf + 7) & ~7, n = nfp >> 3;
    (void)tensor;
    const size_t tile = (size_t)nfp * (chunk_verts + 4) * 4;
    const size_t gram = (size_t)num_pairs(n) * 64 * 8;
    return (tile > gram ? tile : gram) + 128;
}

cudaError_t launch_lm_prep(const DevModel& M, const DevParts& Pt, const LmBuf& a, int batch, cudaStream_t st) {
    lm_prep_kernel<<<batch, 512, lm_prep_smem(M), st>>>(M, Pt, a);
    return cudaGetLastError();
}

// part: 0 = lm_rows_kernel, 1 = lm_gram_kernel, 2 = lm_solve_kernel (one evaluation = the three in order; fp64 path only)
cudaError_t launch_lm_eval_part(const DevModel& M, const DevParts& Pt, const LmBuf& a, int batch, int max_nj, bool tensor,
                                int part, cudaStream_t st) {
    if (part == 0) {
        lm_rows_kernel<<<dim3(a.maxrb, batch), 256, lm_rows_smem(M), st>>>(M, Pt, a);
        return cudaGetLastError();
    }
    if (part == 1) {
        (void)tensor;   // the tensor-core Gram only exists fused into lm_flow_kernel<true>
        const size_t gsm = lm_gram_smem(M, max_nj, a.chunk_verts, false);
        cudaError_t e = cudaFuncSetAttribute(lm_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
        if (e != cudaSuccess) return e;
        lm_gram_kernel<<<dim3(a.maxc, batch), kGramThreads, gsm, st>>>(M, Pt, a);
        return cudaGetLastError();
    }
    const size_t ssm = solve_smem_bytes(M.J, M.K, M.gmmC);
    cudaError_t e = cudaFuncSetAttribute(lm_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
    if (e != cudaSuccess) return e;
    lm_solve_kernel<<<batch, kSolveThreads, ssm, st>>>(M, Pt, a);
    return cudaGetLastError();
}

size_t lm_flow_smem(const DevModel& M, int max_nj, int chunk_verts, bool tensor) {
    size_t b = tensor ? tc_smem_bytes(M.J, M.K) : lm_gram_smem(M, max_nj, chunk_verts, false);
    if (!tensor) {   // (the fused fp64 task keeps the joint tables in front of the tile)
        const size_t fused = fused64_head_bytes(tab_doubles(M.J, M.K), M.K) + lm_gram_smem(M, max_nj, chunk_verts, false);
        b = b > lm_rows_smem(M) ? b : lm_rows_smem(M);
        b = b > fused ? b : fused;
    }
    const size_t ssm = solve_smem_bytes(M.J, M.K, M.gmmC) + (tensor ? 1024 : 0);
    return b > ssm ? b : ssm;
}
template <bool TC, int OCC>
static cudaError_t launch_flow_variant(const DevModel& M, const DevParts& Pt, const LmBuf& a, int ctas, size_t sm, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(lm_flow_kernel<TC, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    lm_flow_kernel<TC, OCC><<<ctas, 256, sm, st>>>(M, Pt, a);
    return cudaGetLastError();
}
cudaError_t launch_lm_flow(const DevModel& M, const DevParts& Pt, const LmBuf& a, int max_nj, int ctas, int occ, cudaStream_t st) {
    const size_t sm = lm_flow_smem(M, max_nj, a.chunk_verts, a.tensor != 0);
    if (a.tensor) {
        if (occ >= 4) return launch_flow_variant<true, 4>(M, Pt, a, ctas, sm, st);
        if (occ == 3) return launch_flow_variant<true, 3>(M, Pt, a, ctas, sm, st);
        return launch_flow_variant<true, 2>(M, Pt, a, ctas, sm, st);
    }
    if (occ >= 3) return launch_flow_variant<false, 3>(M, Pt, a, ctas, sm, st);
    return launch_flow_variant<false, 2>(M, Pt, a, ctas, sm, st);
}
size_t lm_flow_smem_bytes(const DevModel& M, int max_nj, int chunk_verts, bool tensor) { return lm_flow_smem(M, max_nj, chunk_verts, tensor); }
// the tensor path needs every column group to fit the tile (record fields) and the gradient scratch (columns)
bool lm_tensor_supported(int max_nj, int K) { return tc_fields(max_nj, K) <= kTcRows; }


long long lm_part_stride(int max_nj, int J, int K) {   // [ J^T J triangle | J^T r (fp64 path) or P | R | M | T (tensor path) ]
    const int Lg = group_L(max_nj, K), ga = tc_gacc(J, K);
    return (long long)((tri_count(Lg) + (Lg > ga ? Lg : ga) + 1) & ~1);
}
int lm_tab_doubles(int J, int K) { return tab_doubles(J, K); }
int lm_rec_floats(int max_nj, int K) { return rec_floats(max_nj, K); }
int lm_rec_slots(int V) { return (V + 3 * kMaxGroups + 3) & ~3; }
size_t lm_gram_smem_bytes(int max_nj, int K, int chunk_verts, bool tensor) {
    DevModel M{};
    M.K = K;
    return lm_gram_smem(M, max_nj, chunk_verts, tensor);
}

}  // namespace avb
