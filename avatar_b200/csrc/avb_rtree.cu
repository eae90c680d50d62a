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

// ---------------------------------------------------------------------------------------------
// RTree::postProcess (RTree.cpp:3422-3450): suppressPartNonMax (:125-232) / removeSmallPieces (:234-320) + upscaleGrid
// ---------------------------------------------------------------------------------------------
// The reference's flood fill is SEQUENTIAL and, for interval > 1, order dependent: the downward probe tests pixel
// (r + 1, c) but pushes the id of (r + interval, c) unmarked and whatever its label (:176), so components pass through
// grid pixels, erase them with the loser, and those pixels may still seed their own component later.  A parallel
// connected-components labelling cannot reproduce that, so the scan itself is kept as it is -- raster order of the seeds,
// LIFO stack, probes up / down / left / right -- and the parallelism is across the frames of the batch: one warp per
// frame, lane 0 walks, all lanes erase component lists, un-mark and fill the gaps.  The four probes of a popped pixel
// touch four different pixels, so their loads are issued together (one L2 round trip per pixel instead of four).
// Component lists live in an append-only arena (a losing component is reclaimed at once, a winner stays): at most
// 2 x grid pixels ids per frame.
__global__ void __launch_bounds__(32)
rtree_postprocess_kernel(RTreePostArgs a) {
    const int f = blockIdx.x, lane = threadIdx.x;
    uint8_t* img = a.parts + (size_t)f * a.height * a.width;
    const int W = a.width, iv = a.interval;
    int x0 = 0, y0 = 0, x1 = a.width - 1, y1 = a.height - 1;
    if (a.roi) { x0 = a.roi[4 * f]; y0 = a.roi[4 * f + 1]; x1 = a.roi[4 * f + 2]; y1 = a.roi[4 * f + 3]; }
    int* arena = a.arena + (size_t)f * a.cap;
    int* stk = a.stack + (size_t)f * a.cap;
    double* com_pre = a.com_pre + (size_t)f * 2 * a.num_parts;
    const int cap = (int)a.cap;
    constexpr int kMaxP = 64;
    __shared__ int s_best_start[kMaxP], s_best_len[kMaxP];
    __shared__ double s_best_score[kMaxP], s_com_best[2 * kMaxP];
    __shared__ int s_erase_start, s_erase_len, s_state;   // s_state: 0 = scanning, 1 = done, 2 = overflow
    for (int i = lane; i < a.num_parts; i += 32) {
        s_best_start[i] = 0; s_best_len[i] = 0; s_best_score[i] = 0.0;
        s_com_best[2 * i] = 0.0; s_com_best[2 * i + 1] = 0.0;
    }
    if (lane == 0) s_state = 0;
    __syncwarp();
    const int hi_bit = 65536 * iv;
    const size_t thresh = (size_t)(a.height * a.width / (iv * iv) * 0.0005);   // removeSmallPieces: scaledThresh
    int top = 0;                       // arena fill (lane 0)
    int rr = y0, cc = x0;              // scan position (lane 0)
    for (;;) {
        if (lane == 0) {
            s_erase_len = 0;
            // ---- next seed in raster order ----
            bool found = false;
            for (; rr <= y1 && !found; ) {
                for (; cc <= x1; cc += iv) {
                    if (img[(size_t)rr * W + cc] < 128) { found = true; break; }
                }
                if (!found) { rr += iv; cc = x0; }
            }
            if (!found) {
                s_state = 1;
            } else {
                const uint8_t val = img[(size_t)rr * W + cc];
                img[(size_t)rr * W + cc] = (uint8_t)(val + 128);
                const int start = top;
                int sp = 0;
                bool ovf = top + 1 >= cap;
                if (!ovf) {
                    stk[sp++] = (rr << 16) + cc;
                    arena[top++] = (rr << 16) + cc;
                }
                double c0 = 0.0, c1 = 0.0;
                const bool has_prev = a.part_map_type == 0 && com_pre[2 * val] >= 0.;
                while (sp > 0 && !ovf) {
                    const int id = stk[--sp];
                    const int cur_c = id & 0xFFFF, cur_r = id >> 16;
                    const bool pu = cur_r >= y0 + iv, pd = cur_r <= y1 - iv, pl = cur_c >= x0 + iv, pr = cur_c <= x1 - iv;
                    uint8_t* qu = img + (size_t)(cur_r - iv) * W + cur_c;
                    uint8_t* qd = img + (size_t)(cur_r + 1) * W + cur_c;        // the reference probes row r + 1 ...
                    uint8_t* ql = img + (size_t)cur_r * W + cur_c - iv;
                    uint8_t* qr = img + (size_t)cur_r * W + cur_c + iv;
                    const uint8_t vu = pu ? *qu : 255, vd = pd ? *qd : 255, vl = pl ? *ql : 255, vr = pr ? *qr : 255;
                    if (top + 4 >= cap || sp + 4 >= cap) { ovf = true; break; }
                    if (vu == val) { *qu = (uint8_t)(val + 128); arena[top++] = id - hi_bit; stk[sp++] = id - hi_bit; }
                    if (vd == val) { *qd = (uint8_t)(val + 128); arena[top++] = id + hi_bit; stk[sp++] = id + hi_bit; }   // ... and pushes row r + interval
                    if (vl == val) { *ql = (uint8_t)(val + 128); arena[top++] = id - iv; stk[sp++] = id - iv; }
                    if (vr == val) { *qr = (uint8_t)(val + 128); arena[top++] = id + iv; stk[sp++] = id + iv; }
                    c0 += cur_c;
                    c1 += cur_r;
                }
                if (ovf) {
                    s_state = 2;
                } else {
                    const int len = top - start;
                    if (a.part_map_type == 0) {
                        double score = (double)len;
                        c0 /= len;
                        c1 /= len;
                        if (has_prev) {
                            const double d0 = c0 - com_pre[2 * val], d1 = c1 - com_pre[2 * val + 1];
                            score -= __dmul_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), a.dist_w);
                        }
                        if (score > s_best_score[val]) {
                            s_best_score[val] = score;
                            s_com_best[2 * val] = c0;
                            s_com_best[2 * val + 1] = c1;
                            s_erase_start = s_best_start[val];   // the previous best of this part is erased
                            s_erase_len = s_best_len[val];
                            s_best_start[val] = start;
                            s_best_len[val] = len;
                        } else {
                            s_erase_start = start;
                            s_erase_len = len;
                            top = start;                          // reclaim
                        }
                    } else if ((size_t)len < thresh) {
                        s_erase_start = start;
                        s_erase_len = len;
                        top = start;
                    } else {
                        top = start;                              // kept pieces need no list
                    }
                }
                cc += iv;   // the seed itself is marked now; continue the raster scan after it
            }
        }
        __syncwarp();
        if (s_state != 0) break;
        const int es = s_erase_start, el = s_erase_len;
        for (int i = lane; i < el; i += 32) {
            const int id = arena[es + i];
            img[(size_t)(id >> 16) * W + (id & 0xFFFF)] = 255;
        }
        __syncwarp();
    }
    __syncwarp();
    if (s_state == 2) {
        if (lane == 0) a.overflow[f] = 1;
        return;
    }
    if (a.part_map_type == 0)
        for (int i = lane; i < a.num_parts; i += 32) {
            if (s_best_len[i] == 0) {
                com_pre[2 * i] = -1.;
            } else {
                com_pre[2 * i] = s_com_best[2 * i];
                com_pre[2 * i + 1] = s_com_best[2 * i + 1];
            }
        }
    // un-mark (:205-227)
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    if (bw > 0 && bh > 0)
        for (long long p = lane; p < (long long)bw * bh; p += 32) {
            uint8_t* q = img + (size_t)(y0 + (int)(p / bw)) * W + x0 + (int)(p % bw);
            const uint8_t v = *q;
            if (v >= 128 && v != 255) *q = (uint8_t)(v - 128);
        }
    __syncwarp();
    // upscaleGrid (:70-100): every interval x interval cell takes the label of its grid pixel (clamped to the image)
    if (iv > 1 && bw > 0)
        for (int rg = y0 + iv; rg <= y1; rg += iv) {
            const int ncell = (x1 - x0) / iv + 1;
            for (int p = lane; p < ncell * iv; p += 32) {
                const int cell = p / iv, dr = p % iv, r = rg + dr;
                if (r > y1) continue;
                const int cx = x0 + cell * iv;
                const uint8_t v = img[(size_t)rg * W + cx];
                for (int k = (dr == 0 ? 1 : 0); k < iv && cx + k < W; ++k) img[(size_t)r * W + cx + k] = v;
            }
            __syncwarp();
        }
}

cudaError_t launch_rtree_postprocess(const RTreePostArgs& a, int batch, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    rtree_postprocess_kernel<<<batch, 32, 0, st>>>(a);
    return cudaGetLastError();
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
