// avb_render.cu -- the reference's painter's-algorithm renderer on the device:
// AvatarRenderer::renderDepth / renderPartMask / renderFaces (AvatarRenderer.cpp:72-98, 170-216) with
// getProjectedPoints (:11-24) and getOrderedFaces (:41-70).
//
// The reference paints 13 776 triangles one after the other, far to near, so that the last painter of a pixel wins.
// Here the order becomes a rank and the painting an atomicMax (avb_paint.h has the coverage and value functions, shared
// with the CPU check of tests/cpp/paint_check.cpp):
//   render_prepare_kernel  one CTA per frame: project the vertices (double maths, float result), build the 64-bit keys
//                          (mean z as float, descending; face index ascending) and sort them with a bitonic network in
//                          shared memory (16 384 keys = 128 KB) -> faces in paint order.
//   render_cover_kernel    one thread per face: enumerate exactly the pixels the reference painter writes for this
//                          face and atomicMax its rank into the winner image(s).
//   render_resolve_kernel  one thread per pixel: the value the winning face's painter writes there.
// Results are bit-identical to the sequential painter (oracle/render_oracle.cpp), with ascending face index among equal
// keys where the reference's std::sort leaves the order unspecified.
#include "avb_device.cuh"
#include "avb_kernels.h"
#include "avb_paint.h"

namespace avb {

using namespace paint;

constexpr int kSortN = 16384;          // keys per frame in shared memory (F <= 16384)
constexpr int kPrepThreads = 1024;

__global__ void __launch_bounds__(kPrepThreads, 1)
render_prepare_kernel(RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x;
    const double* cloud = a.cloud + (size_t)f * 3 * a.V;
    P2* proj = reinterpret_cast<P2*>(a.proj) + (size_t)f * a.V;
    for (int v = tid; v < a.V; v += kPrepThreads)
        proj[v] = project(cloud[3 * (size_t)v], cloud[3 * (size_t)v + 1], cloud[3 * (size_t)v + 2], a.fx, a.cx, a.fy, a.cy);
    for (int i = tid; i < kSortN; i += kPrepThreads) {
        unsigned long long k = 0xFFFFFFFFFFFFFFFFull;   // padding sorts last
        if (i < a.F) {
            const int* t = a.faces + 3 * (size_t)i;
            k = order_key(face_key(cloud[3 * (size_t)t[0] + 2], cloud[3 * (size_t)t[1] + 2], cloud[3 * (size_t)t[2] + 2]), i);
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= kSortN; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < kSortN / 2; t += kPrepThreads) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = keys[lo], y = keys[hi];
                if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    int* order = a.order + (size_t)f * a.F;
    for (int i = tid; i < a.F; i += kPrepThreads) {
        const int face = (int)(keys[i] & 0xFFFFFFFFull);
        order[i] = face;
        if (a.rank_of) a.rank_of[(size_t)f * a.F + face] = i;
    }
}

__device__ __forceinline__ RenderView view_of(const RenderArgs& a, int f) {
    RenderView v;
    v.cloud = a.cloud + (size_t)f * 3 * a.V;
    v.faces = a.faces;
    v.proj = reinterpret_cast<const P2*>(a.proj) + (size_t)f * a.V;
    v.vpart = a.vpart;
    v.W = a.width;
    v.H = a.height;
    return v;
}

__global__ void __launch_bounds__(128)
render_cover_kernel(RenderArgs a) {
    const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.F) return;
    const RenderView v = view_of(a, f);
    const size_t px = (size_t)a.width * a.height;
    face_cover(v, a.order[(size_t)f * a.F + i], (unsigned)i + 1u, a.win_depth ? a.win_depth + f * px : nullptr,
               a.win_parts ? a.win_parts + f * px : nullptr, a.win_faces ? a.win_faces + f * px : nullptr,
               [](unsigned* p, unsigned r) { atomicMax(p, r); });
}

__global__ void __launch_bounds__(256)
render_resolve_kernel(RenderArgs a) {
    const int f = blockIdx.y;
    const size_t px = (size_t)a.width * a.height;
    const unsigned p = blockIdx.x * 256u + threadIdx.x;
    if (p >= px) return;
    const int i = (int)(p / (unsigned)a.width), j = (int)(p - (unsigned)i * (unsigned)a.width);
    const RenderView v = view_of(a, f);
    const int* order = a.order + (size_t)f * a.F;
    if (a.depth_out) a.depth_out[f * px + p] = resolve_depth(v, order, a.win_depth[f * px + p], i, j);
    if (a.parts_out) a.parts_out[f * px + p] = resolve_parts(v, order, a.win_parts[f * px + p], i, j);
    if (a.faces_out) a.faces_out[f * px + p] = resolve_faces(a.win_faces[f * px + p]);
}

// ---- renderLambert (AvatarRenderer.cpp:103-172) ----
// The reference adds the unit normal of every face to its three vertices while it walks the faces in PAINT order, so the
// floating-point sum at a vertex depends on the frame's face order.  One thread per vertex: gather the incident faces
// (static CSR), order them by the frame's paint position, add their normals in that order, normalise, flip towards the
// camera, light (paint::vertex_lambert).
constexpr int kMaxValence = 32;

__global__ void __launch_bounds__(128)
render_vertex_lambert_kernel(RenderArgs a) {
    const int f = blockIdx.y, v = blockIdx.x * 128 + threadIdx.x;
    if (v >= a.V) return;
    const double* cloud = a.cloud + (size_t)f * 3 * a.V;
    const int* rank_of = a.rank_of + (size_t)f * a.F;
    const int s0 = a.vf_start[v], n = min(a.vf_start[v + 1] - s0, kMaxValence);
    int fr[kMaxValence], fc[kMaxValence];
    for (int i = 0; i < n; ++i) {   // insertion sort by paint position
        const int face = a.vf_list[s0 + i], r = rank_of[face];
        int k = i;
        while (k > 0 && fr[k - 1] > r) {
            fr[k] = fr[k - 1];
            fc[k] = fc[k - 1];
            --k;
        }
        fr[k] = r;
        fc[k] = face;
    }
    double ns[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
        const int* t = a.faces + 3 * (size_t)fc[i];
        double nn[3];
        face_unit_normal(cloud + 3 * (size_t)t[0], cloud + 3 * (size_t)t[1], cloud + 3 * (size_t)t[2], nn);
        // a degenerate face listing the vertex twice adds its normal twice, as the reference's loop over j does
        for (int j = 0; j < 3; ++j)
            if (t[j] == v) { ns[0] = dadd(ns[0], nn[0]); ns[1] = dadd(ns[1], nn[1]); ns[2] = dadd(ns[2], nn[2]); }
    }
    a.vlam[(size_t)f * a.V + v] = vertex_lambert(cloud + 3 * (size_t)v, ns);
}

__global__ void __launch_bounds__(128)
render_cover_lambert_kernel(RenderArgs a) {
    const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.F) return;
    const RenderView v = view_of(a, f);
    const size_t px = (size_t)a.width * a.height;
    face_cover_lambert(v, a.order[(size_t)f * a.F + i], (unsigned)i + 1u, a.win_lambert + f * px, [](unsigned* p, unsigned r) { atomicMax(p, r); });
}

__global__ void __launch_bounds__(256)
render_resolve_lambert_kernel(RenderArgs a) {
    const int f = blockIdx.y;
    const size_t px = (size_t)a.width * a.height;
    const unsigned p = blockIdx.x * 256u + threadIdx.x;
    if (p >= px) return;
    const int i = (int)(p / (unsigned)a.width), j = (int)(p - (unsigned)i * (unsigned)a.width);
    const RenderView v = view_of(a, f);
    a.lambert_out[f * px + p] = resolve_lambert(v, a.order + (size_t)f * a.F, a.vlam + (size_t)f * a.V, a.win_lambert[f * px + p], i, j);
}

int render_max_valence() { return kMaxValence; }

cudaError_t launch_render_lambert(const RenderArgs& a, int batch, cudaStream_t st) {
    const size_t smem = (size_t)kSortN * 8;
    cudaError_t e = cudaFuncSetAttribute(render_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    render_prepare_kernel<<<batch, kPrepThreads, smem, st>>>(a);
    render_vertex_lambert_kernel<<<dim3((a.V + 127) / 128, batch), 128, 0, st>>>(a);
    render_cover_lambert_kernel<<<dim3((a.F + 127) / 128, batch), 128, 0, st>>>(a);
    const size_t px = (size_t)a.width * a.height;
    render_resolve_lambert_kernel<<<dim3((unsigned)((px + 255) / 256), batch), 256, 0, st>>>(a);
    return cudaGetLastError();
}

int render_max_faces() { return kSortN; }

cudaError_t launch_render(const RenderArgs& a, int batch, cudaStream_t st, cudaEvent_t* ev4) {
    const size_t smem = (size_t)kSortN * 8;
    cudaError_t e = cudaFuncSetAttribute(render_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (ev4) cudaEventRecord(ev4[0], st);
    render_prepare_kernel<<<batch, kPrepThreads, smem, st>>>(a);
    if (ev4) cudaEventRecord(ev4[1], st);
    render_cover_kernel<<<dim3((a.F + 127) / 128, batch), 128, 0, st>>>(a);
    if (ev4) cudaEventRecord(ev4[2], st);
    const size_t px = (size_t)a.width * a.height;
    render_resolve_kernel<<<dim3((unsigned)((px + 255) / 256), batch), 256, 0, st>>>(a);
    if (ev4) cudaEventRecord(ev4[3], st);
    return cudaGetLastError();
}

}  // namespace avb
