// avb_synth.cpp -- synthetic-data harness (host CPU code; NOT on the fit path, NOT part of libavatar_b200.so).
// Functional stand-in for AvatarRenderer::renderDepth / renderPartMask (AvatarRenderer.cpp:72-216,
// AvatarHelpers.cpp:61-313) and for the depth -> cloud back-projection the reference's callers do
// (optim.cpp:104-120, demo.cpp:226-250, Calibration.cpp:68-74).  Own z-buffer rasteriser; exactness
// against the reference's painter's algorithm is not required.
#include <cstdint>
#define AVB_ERR_INVALID 1
#define AVB_OK 0

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

extern "C" {

int avb_synth_render(const double* cloud, int32_t V, const int32_t* faces, int32_t F, const int32_t* vertex_part,
                     int32_t width, int32_t height, float fx, float fy, float cx, float cy, float* depth_out,
                     uint8_t* part_out) {
    if (!cloud || !faces || !depth_out || width <= 0 || height <= 0) return AVB_ERR_INVALID;
    const size_t npx = (size_t)width * height;
    std::fill(depth_out, depth_out + npx, 0.f);
    if (part_out) std::fill(part_out, part_out + npx, (uint8_t)255);
    std::vector<double> px(V), py(V);
    for (int v = 0; v < V; ++v) {  // AvatarRenderer::getProjectedPoints: y is flipped
        const double z = cloud[3 * (size_t)v + 2];
        px[v] = cloud[3 * (size_t)v] * fx / z + cx;
        py[v] = -cloud[3 * (size_t)v + 1] * fy / z + cy;
    }
    for (int f = 0; f < F; ++f) {
        const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
        const double za = cloud[3 * (size_t)ia + 2], zb = cloud[3 * (size_t)ib + 2], zc = cloud[3 * (size_t)ic + 2];
        if (za <= 0 || zb <= 0 || zc <= 0) continue;
        const double ax = px[ia], ay = py[ia], bx = px[ib], by = py[ib], cx2 = px[ic], cy2 = py[ic];
        const double area = (bx - ax) * (cy2 - ay) - (by - ay) * (cx2 - ax);
        if (std::fabs(area) < 1e-12) continue;
        const int x0 = std::max(0, (int)std::ceil(std::min({ax, bx, cx2})));
        const int x1 = std::min(width - 1, (int)std::floor(std::max({ax, bx, cx2})));
        const int y0 = std::max(0, (int)std::ceil(std::min({ay, by, cy2})));
        const int y1 = std::min(height - 1, (int)std::floor(std::max({ay, by, cy2})));
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const double w0 = ((bx - x) * (cy2 - y) - (by - y) * (cx2 - x)) / area;
                const double w1 = ((cx2 - x) * (ay - y) - (cy2 - y) * (ax - x)) / area;
                const double w2 = 1.0 - w0 - w1;
                if (w0 < 0 || w1 < 0 || w2 < 0) continue;
                const float z = (float)(w0 * za + w1 * zb + w2 * zc);
                float& d = depth_out[(size_t)y * width + x];
                if (d == 0.f || z < d) {
                    d = z;
                    if (part_out && vertex_part) {  // part of the nearest projected vertex of the hit triangle
                        const double da = (ax - x) * (ax - x) + (ay - y) * (ay - y);
                        const double db = (bx - x) * (bx - x) + (by - y) * (by - y);
                        const double dc = (cx2 - x) * (cx2 - x) + (cy2 - y) * (cy2 - y);
                        const int best = (da <= db && da <= dc) ? ia : (db <= dc ? ib : ic);
                        part_out[(size_t)y * width + x] = (uint8_t)vertex_part[best];
                    }
                }
            }
    }
    return AVB_OK;
}

int64_t avb_synth_backproject(const float* depth, const uint8_t* part, int32_t width, int32_t height, float fx,
                              float fy, float cx, float cy, int32_t interval, double* cloud_out,
                              int32_t* labels_out, int64_t max_points) {
    if (!depth || !cloud_out || interval <= 0) return -1;
    int64_t n = 0;
    for (int r = 0; r < height; r += interval)
        for (int c = 0; c < width; c += interval) {
            const float z = depth[(size_t)r * width + c];
            if (z <= 0.f) continue;
            if (n >= max_points) return -2;
            // CameraIntrin::to3D in float (Calibration.cpp:68-74), then y negated (optim.cpp:115-120)
            const float X = ((float)c - cx) * z / fx;
            const float Y = ((float)r - cy) * z / fy;
            cloud_out[3 * n] = (double)X;
            cloud_out[3 * n + 1] = -(double)Y;
            cloud_out[3 * n + 2] = (double)z;
            if (labels_out) labels_out[n] = part ? (int32_t)part[(size_t)r * width + c] : 0;
            ++n;
        }
    return n;
}

}  // extern "C"
