// avb_cloud.cu -- data-cloud construction on the device: depth image + body-part
// label image -> the 3xN fp64 data cloud and the N part labels that AvatarOptimizer::optimize consumes.
//
// Replaces the host loops of demo.cpp:215-250 (count the foreground pixels of the background-subtractor bounding
// box at stride `interval`, then fill dataCloud / dataPartLabels in raster order, y negated) together with
// CameraIntrin::depthToXYZ (Calibration.cpp:83-95: x = (c - cx) z / fx, y = (r - cy) z / fy in float).
//
//   cloud_count_kernel    grid = (strips, frames): foreground pixels per strip of 8 sampled rows, label check.
//   cloud_compact_kernel  grid = (strips, frames): stable (raster-order) compaction of the strip into the frame's
//                         slice of the batch cloud: one warp per sampled row, ballot/popc ranks, float maths with
//                         explicit round-to-nearest operations (no FMA contraction, IEEE division) so that every
//                         coordinate is bit-identical to the CPU loop.
// Both are streaming kernels bound by HBM: the label image is read once per kernel (the second read of
// cloud_compact_kernel inside the CTA hits L1), depth only where a pixel is foreground; 28 B written per point.
#include "avb_device.cuh"
#include "avb_kernels.h"

namespace avb {

constexpr int kStripRows = 8;       // sampled rows per CTA = warps per CTA
constexpr int kCloudThreads = 32 * kStripRows;

__device__ __forceinline__ void roi_of(const CloudArgs& a, int f, int& x0, int& y0, int& ncols, int& nrows) {
    int x1, y1;
    if (a.roi) {
        x0 = max(a.roi[4 * f], 0);
        y0 = max(a.roi[4 * f + 1], 0);
        x1 = min(a.roi[4 * f + 2], a.width - 1);
        y1 = min(a.roi[4 * f + 3], a.height - 1);
    } else {
        x0 = y0 = 0;
        x1 = a.width - 1;
        y1 = a.height - 1;
    }
    ncols = (x1 >= x0) ? (x1 - x0) / a.interval + 1 : 0;
    nrows = (y1 >= y0) ? (y1 - y0) / a.interval + 1 : 0;
}

// foreground flags of the four sampled pixels [4k, 4k+4) of one sampled row; labels in lab[4] (255 = background)
__device__ __forceinline__ unsigned load4(const uint8_t* row, int x0, int interval, int ncols, int k, unsigned lab[4]) {
    const int s0 = 4 * k;
    unsigned bits = 0;
    if (interval == 1 && s0 + 4 <= ncols && ((reinterpret_cast<uintptr_t>(row + x0 + s0) & 3) == 0)) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(row + x0 + s0);
        lab[0] = v.x; lab[1] = v.y; lab[2] = v.z; lab[3] = v.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= (lab[j] != 255u) ? (1u << j) : 0u;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lab[j] = 255u;
            if (s0 + j < ncols) lab[j] = row[x0 + (s0 + j) * interval];
            bits |= (lab[j] != 255u) ? (1u << j) : 0u;
        }
    }
    return bits;
}

__global__ void __launch_bounds__(kCloudThreads)
cloud_count_kernel(CloudArgs a) {
    __shared__ int s_cnt[kStripRows];
    const int f = blockIdx.y, strip = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x0, y0, ncols, nrows;
    roi_of(a, f, x0, y0, ncols, nrows);
    const int sr = strip * kStripRows + wid;   // sampled row of this warp
    int cnt = 0;
    bool bad = false;
    if (sr < nrows) {
        const uint8_t* row = a.parts + ((size_t)f * a.height + (y0 + sr * a.interval)) * a.width;
        for (int k = lane; 4 * k < ncols; k += 32) {
            unsigned lab[4];
            const unsigned bits = load4(row, x0, a.interval, ncols, k, lab);
            cnt += __popc(bits);
#pragma unroll
            for (int j = 0; j < 4; ++j) bad |= (lab[j] != 255u && (int)lab[j] >= a.num_parts);
        }
    }
    cnt = warp_sum_i(cnt);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.bad_label + f, 1);   // the reference exits (demo.cpp:232-239)
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kStripRows; ++w) t += s_cnt[w];
        a.strip_count[(size_t)f * a.max_strips + strip] = t;
    }
}

__global__ void __launch_bounds__(kCloudThreads)
cloud_compact_kernel(CloudArgs a) {
    __shared__ int s_cnt[kStripRows];
    const int f = blockIdx.y, strip = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x0, y0, ncols, nrows;
    roi_of(a, f, x0, y0, ncols, nrows);
    const int sr = strip * kStripRows + wid;
    const int r = y0 + sr * a.interval;
    const uint8_t* row = a.parts + ((size_t)f * a.height + r) * a.width;
    // pass 1: foreground pixels of this warp's row (re-read below: L1 hit)
    int cnt = 0;
    if (sr < nrows)
        for (int k = lane; 4 * k < ncols; k += 32) {
            unsigned lab[4];
            cnt += __popc(load4(row, x0, a.interval, ncols, k, lab));
        }
    cnt = warp_sum_i(cnt);
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    if (sr >= nrows) return;
    long long base = a.strip_offset[(size_t)f * a.max_strips + strip];   // first point of the strip in the batch cloud
    for (int w = 0; w < wid; ++w) base += s_cnt[w];
    // pass 2: raster-order ranks inside the row, 128 sampled pixels per step
    const float* drow = a.depth + ((size_t)f * a.height + r) * a.width;
    const float ry = __fsub_rn((float)r, a.cy);
    for (int k0 = 0; 4 * k0 < ncols; k0 += 32) {
        const int k = k0 + lane;
        unsigned lab[4];
        const unsigned bits = (4 * k < ncols) ? load4(row, x0, a.interval, ncols, k, lab) : 0u;
        const int mine = __popc(bits);
        int incl = mine;   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        long long pos = base + (incl - mine);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (bits & (1u << j)) {
                const int c = x0 + (4 * k + j) * a.interval;
                const float z = drow[c];
                // Calibration.cpp:91: Vec3f((c - cx) * z / fx, (r - cy) * z / fy, z); demo.cpp:244-246 negates y
                const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)c, a.cx), z), a.fx);
                const float y = __fdiv_rn(__fmul_rn(ry, z), a.fy);
                double* o = a.cloud + 3 * pos;
                o[0] = (double)x;
                o[1] = (double)(-y);
                o[2] = (double)z;
                a.labels[pos] = (int)lab[j];
                ++pos;
            }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

int cloud_strip_rows() { return kStripRows; }

cudaError_t launch_cloud_count(const CloudArgs& a, int strips, int batch, cudaStream_t st) {
    if (strips > 0 && batch > 0) cloud_count_kernel<<<dim3(strips, batch), kCloudThreads, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_cloud_compact(const CloudArgs& a, int strips, int batch, cudaStream_t st) {
    if (strips > 0 && batch > 0) cloud_compact_kernel<<<dim3(strips, batch), kCloudThreads, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace avb
