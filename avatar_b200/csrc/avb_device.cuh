// avb_device.cuh -- device-side data layout and small math helpers shared by the kernels.
// B200 (sm_100a) only.  See DESIGN.md for the HBM layout and the per-kernel rooflines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace avb {

constexpr int kMaxJ = 32;
constexpr int kMaxK = 16;
constexpr int kMaxParts = 64;
constexpr int kMaxGroups = 16;
constexpr double kFixScale = 68719476736.0;           // 2^36: fixed-point scale of the per-vertex data sums
constexpr double kFixInv = 1.0 / 68719476736.0;
constexpr double kCoordLimit = 32.0;                   // |coordinate| < 32 m so 2^21 points cannot overflow int64
constexpr int kQBlock = 256;                           // granularity of the deterministic sum |d|^2 partials

// Immutable model, device pointers (ark::AvatarModel data, include/Avatar.h:64-151)
struct DevModel {
    int V, J, K, F, P, nx, max_depth;
    const double* vt;          // [3V]        baseCloud
    const float* sd;           // [V][3][K]   keyClouds rows of vertex v (fp32 storage, fp64 maths)
    const float* sdT;          // [3][K][V]   the same, component-major (coalesced when one thread owns one vertex)
    const double* sk_w;        // [V][4]      assignedJoints weights (desc), 0 padded
    const uint8_t* sk_j;       // [V][4]      assignedJoints joints
    const uint8_t* sk_n;       // [V]
    const uint32_t* anc_mask;  // [J]         bit a set <=> a is j or an ancestor of j
    const int* parent;         // [J]
    const int* depth;          // [J]
    const int* lvl_start;      // [max_depth+2] into lvl_joint
    const int* lvl_joint;      // [J]         joints sorted by (depth, id)
    const double* jbase;       // [3J]        jointShapeRegBase
    const double* jreg;        // [3J][K]     jointShapeReg  (S_j = rows 3j..3j+2)
    const double* Sp;          // [J][3][K]   S_j - S_parent(j)   (AvatarOptimizer.cpp:240-243)
    const int* faces;          // [3F]
    int gmmC, gmmD;
    const double* gmm_mean;    // [C][D]
    const double* gmm_prec;    // [C][D][D]   Sigma^-1 = prec_cho prec_cho^T (full symmetric)
    const double* gmm_clog;    // [C]         consts_log
    const double* gmm_pfull;   // [C][P(P+1)/2] Sigma^-1 zero padded to the tangent layout (rows/cols 6..6+D), packed lower triangle
};

// Per-optimizer part tables (AvatarOptimizer.cpp:1213-1244) and the static column-group schedule
struct DevParts {
    int numParts;
    const int* part_start;     // [numParts+1] into part_verts
    const int* part_verts;     // [V] vertex ids grouped by part, ascending inside a part
    const int* first_part_at;  // [V+1] smallest part p with part_start[p] == i, else -1
    int numGroups;
    const int* gorder;         // [V] vertex ids sorted by (group, id)
    const int* gvstart;        // [numGroups+1] into gorder
    const int* gjoints;        // [numGroups][kMaxJ] joint ids of the group's column set (ascending)
    const int* gnj;            // [numGroups]
    const int* gdoff;          // [numGroups+1] into gdest
    const int* gdest;          // per group, per entry of a chunk partial [ J^T J triangle | J^T r ]: where the solve adds it --
                               // index into the packed P x P lower triangle, tri(P) + column for the gradient, -1 = unused slot
};

struct FrameStats {  // mirrors avb_stats
    int num_points, num_correspondences, num_matched_vertices, iterations, accepted_steps, status;
    double initial_cost, final_cost;
};

#ifdef __CUDACC__
// Eigen Quaterniond::toRotationMatrix, q = (x,y,z,w); R row-major
__device__ __forceinline__ void quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic CTA-wide sum (fixed shuffle tree, then warp partials in warp order). scratch >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double s = 0;
#pragma unroll 1
    for (int i = 0; i < nw; ++i) s += scratch[i];
    return s;
}

#endif  // __CUDACC__

}  // namespace avb
