// avb_rtree.cu -- body-part label prediction on the device: the image form of
// RTree::predictBest (RTree.cpp:3184-3262) and its gap filling (upscaleGrid, RTree.cpp:70-100).
//
// One thread walks the decision tree for one sampled pixel: at every internal node the two probe offsets u, v are
// divided by the pixel's depth, rounded half away from zero (std::round), added to the pixel, looked up inside the
// bounding box (outside it, or at depth 0: BACKGROUND_DEPTH = 20 m, RTree.cpp:325) and the difference is compared with
// the node's threshold.  The leaf's best part (leafBestMatch, RTree.cpp:3455-3463) is the label.  Float division,
// rounding and subtraction are the reference's, so the labels are bit-exact.  Nodes are packed into 32-byte records
// (one sector per visit); the walk is a chain of dependent gathers, i.e. bound by L2 latency / bandwidth, and hides it
// with one thread per pixel at full occupancy.
//
// Reference quirk kept on purpose: the row loop starts with `r = (row += interval)`, so the first row of the bounding
// box is never predicted (RTree.cpp:3196-3199).
#include "avb_device.cuh"
#include "avb_kernels.h"

namespace avb {

constexpr float kBackgroundDepth = 20.f;   // RTree::BACKGROUND_DEPTH (RTree.cpp:325)

__device__ __forceinline__ void rtree_roi(const RTreeArgs& a, int f, int& x0, int& y0, int& x1, int& y1) {
    if (a.roi) {
        x0 = a.roi[4 * f]; y0 = a.roi[4 * f + 1]; x1 = a.roi[4 * f + 2]; y1 = a.roi[4 * f + 3];
    } else {
        x0 = y0 = 0;
        x1 = a.width - 1;   // bot_right == (-1, -1): whole image (RTree.cpp:3190-3193)
        y1 = a.height - 1;
    }
}

// grid = (ceil(sampled pixels of the largest box / 2048), frames), grid-stride over the box
__global__ void __launch_bounds__(256)
rtree_predict_kernel(RTreeArgs a) {
    const int f = blockIdx.y;
    int x0, y0, x1, y1;
    rtree_roi(a, f, x0, y0, x1, y1);
    const int ncols = (x1 >= x0) ? (x1 - x0) / a.interval + 1 : 0;
    const int nrows = (y1 - y0) / a.interval;             // rows y0 + interval, y0 + 2 interval, ... <= y1
    if (ncols <= 0 || nrows <= 0) return;
    const float* depth = a.depth + (size_t)f * a.height * a.width;
    const unsigned total = (unsigned)ncols * (unsigned)nrows;   // < 2^31 pixels per image: 32-bit index maths
    // grid-stride: most pixels are background (one depth load, no walk); eight per thread keep the launch small
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned q = i / (unsigned)ncols;
        const int r = y0 + (int)(q + 1) * a.interval, c = x0 + (int)(i - q * (unsigned)ncols) * a.interval;
        if (r < 0 || r >= a.height || c < 0 || c >= a.width) continue;   // a box outside the image is undefined in the reference
        const float sampleDepth = depth[(size_t)r * a.width + c];
        if (sampleDepth == 0.f) continue;
        int nodeid = 0;
        RTreeNode nd = a.nodes[0];
        while (nd.leafid == -1) {
            // Eigen: Vector2f / float is a component-wise IEEE division; std::round is half away from zero
            const int ux = (int)roundf(__fdiv_rn(nd.ux, sampleDepth)) + c, uy = (int)roundf(__fdiv_rn(nd.uy, sampleDepth)) + r;
            const int vx = (int)roundf(__fdiv_rn(nd.vx, sampleDepth)) + c, vy = (int)roundf(__fdiv_rn(nd.vy, sampleDepth)) + r;
            float zu = kBackgroundDepth, zv = kBackgroundDepth;
            // inside the box AND inside the image (the entry points reject boxes that leave the image; this keeps the
            // gather in bounds even so)
            if (!(ux < x0 || uy < y0 || ux > x1 || uy > y1) && (unsigned)ux < (unsigned)a.width && (unsigned)uy < (unsigned)a.height) {
                zu = depth[(size_t)uy * a.width + ux];
                if (zu == 0.f) zu = kBackgroundDepth;
            }
            if (!(vx < x0 || vy < y0 || vx > x1 || vy > y1) && (unsigned)vx < (unsigned)a.width && (unsigned)vy < (unsigned)a.height) {
                zv = depth[(size_t)vy * a.width + vx];
                if (zv == 0.f) zv = kBackgroundDepth;
            }
            nodeid = (__fsub_rn(zu, zv) < nd.thresh) ? nd.lnode : nd.rnode;
            nd = a.nodes[nodeid];
        }
        a.parts[((size_t)f * a.height + r) * a.width + c] = a.leaf_best[nd.leafid];
    }
}

// upscaleGrid (RTree.cpp:70-100): every interval x interval cell below / right of a predicted pixel takes its label.
// grid = (ceil(box pixels / 2048), frames), grid-stride; the predicted pixels themselves are left alone (no read/write overlap).
__global__ void __launch_bounds__(256)
rtree_upscale_kernel(RTreeArgs a) {
    const int f = blockIdx.y;
    int x0, y0, x1, y1;
    rtree_roi(a, f, x0, y0, x1, y1);
    const int ncols = (x1 >= x0) ? (x1 - x0) / a.interval + 1 : 0;   // cells per row
    const int wcols = ncols * a.interval;                             // memset(ptr + cc, val, interval) may pass x1
    const int hrows = y1 - (y0 + a.interval) + 1;                     // rows y0 + interval .. y1
    if (wcols <= 0 || hrows <= 0) return;
    uint8_t* img = a.parts + (size_t)f * a.height * a.width;
    const unsigned total = (unsigned)wcols * (unsigned)hrows;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const int dr = (int)(i / (unsigned)wcols), dc = (int)(i - (unsigned)dr * (unsigned)wcols);
        const int r = y0 + a.interval + dr, x = x0 + dc;
        const int rr = y0 + a.interval + (dr / a.interval) * a.interval, cc = x0 + (dc / a.interval) * a.interval;
        if (r < 0 || r >= a.height || x < 0 || x >= a.width || (r == rr && x == cc)) continue;   // clamped to the image
        img[(size_t)r * a.width + x] = img[(size_t)rr * a.width + cc];
    }
}

cudaError_t launch_rtree_predict(const RTreeArgs& a, int batch, int max_box_pixels, cudaStream_t st) {
    if (batch <= 0 || max_box_pixels <= 0) return cudaSuccess;
    const int blocks = (max_box_pixels + 2047) / 2048;   // eight pixels per thread
    rtree_predict_kernel<<<dim3(blocks, batch), 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_rtree_upscale(const RTreeArgs& a, int batch, int max_box_pixels, cudaStream_t st) {
    if (batch <= 0 || max_box_pixels <= 0) return cudaSuccess;
    const int blocks = (max_box_pixels + 2047) / 2048;
    rtree_upscale_kernel<<<dim3(blocks, batch), 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace avb
