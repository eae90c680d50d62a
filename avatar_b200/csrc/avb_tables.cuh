// avb_tables.cuh -- per-frame joint tables (forward kinematics + shape tables), built CTA-wide in shared memory.
#pragma once
#include "avb_device.cuh"

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
__device__ inline void build_tables(const DevModel& M, const double* xs, Tables T, bool with_shape) {
    const int J = M.J, K = M.K, tid = threadIdx.x, nt = blockDim.x;
    const double* w = xs + 3 + 4 * J;
#pragma unroll 1
    for (int i = tid; i < 3 * J; i += nt) {
        double s = 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) s += M.jreg[i * K + k] * w[k];
        T.Jr[i] = M.jbase[i] + s;
    }
#pragma unroll 1
    for (int j = tid; j < J; j += nt) quat_to_rot(xs + 3 + 4 * j, T.Rq + 9 * j);
    __syncthreads();
    // one sweep over the kinematic tree, level by level (joints of a level listed in M.lvl_joint): 64 threads per joint,
    // thread e < 9: entry of G_j = G_parent R_j, 9 <= e < 12: pos_j, 12 <= e < 12 + 3K: row of H[j] (needs G_parent only);
    // one CTA barrier per level instead of two sweeps of max_depth barriers each (3 K <= 48: AVB_MAX_SHAPE_KEYS = 16)
    const int per = 3 * K, nitems = 12 + (with_shape ? per : 0);
#pragma unroll 1
    for (int d = 0; d <= M.max_depth; ++d) {
        const int l0 = M.lvl_start[d], l1 = M.lvl_start[d + 1];
#pragma unroll 1
        for (int jj = l0 + (tid >> 6); jj < l1; jj += nt >> 6) {
            const int e = tid & 63;
            if (e >= nitems) continue;
            const int j = M.lvl_joint[jj], pa = M.parent[j];
            const double* R = T.Rq + 9 * j;
            if (pa < 0) {
                if (e < 9) T.G[9 * j + e] = R[e];
                else if (e < 12) T.pos[3 * j + (e - 9)] = xs[e - 9];
                else T.Hj[j * per + (e - 12)] = 0.0;
            } else {
                const double* Gp = T.G + 9 * pa;
                if (e < 9) {
                    const int r = e / 3, c = e - 3 * r;
                    T.G[9 * j + e] = Gp[3 * r] * R[c] + Gp[3 * r + 1] * R[3 + c] + Gp[3 * r + 2] * R[6 + c];
                } else if (e < 12) {
                    const int r = e - 9;
                    const double v0 = T.Jr[3 * j] - T.Jr[3 * pa], v1 = T.Jr[3 * j + 1] - T.Jr[3 * pa + 1],
                                 v2 = T.Jr[3 * j + 2] - T.Jr[3 * pa + 2];
                    T.pos[3 * j + r] = Gp[3 * r] * v0 + Gp[3 * r + 1] * v1 + Gp[3 * r + 2] * v2 + T.pos[3 * pa + r];
                } else {
                    const int i = e - 12, r = i / K, m = i - r * K;
                    const double* sp = M.Sp + (size_t)j * per;
                    T.Hj[j * per + i] = Gp[3 * r] * sp[m] + Gp[3 * r + 1] * sp[K + m] + Gp[3 * r + 2] * sp[2 * K + m] +
                                        T.Hj[pa * per + r * K + m];
                }
            }
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int j = tid; j < J; j += nt) {
        const double* Gj = T.G + 9 * j;
#pragma unroll 1
        for (int r = 0; r < 3; ++r)
            T.tau[3 * j + r] = T.pos[3 * j + r] -
                               (Gj[3 * r] * T.Jr[3 * j] + Gj[3 * r + 1] * T.Jr[3 * j + 1] + Gj[3 * r + 2] * T.Jr[3 * j + 2]);
    }
    if (with_shape) {
#pragma unroll 1
        for (int i = tid; i < J * per; i += nt) {
            const int j = i / per, r = (i % per) / K, m = i % K;
            const double* Gj = T.G + 9 * j;
            const double* S = M.jreg + (size_t)3 * j * K;
            T.C[i] = T.Hj[i] - (Gj[3 * r] * S[m] + Gj[3 * r + 1] * S[K + m] + Gj[3 * r + 2] * S[2 * K + m]);
        }
    }
    __syncthreads();
}

}  // namespace avb
