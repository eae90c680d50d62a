// tests/cpp/paint_check.cpp -- runs the rank-form renderer of avatar_b200/csrc/avb_paint.h on the CPU (the same
// __host__ __device__ code the CUDA kernels call) so that tests/test_oracle.py can compare it with the sequential
// painter of oracle/render_oracle.cpp without a GPU.  Compiled with -ffp-contract=off.
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../../avatar_b200/csrc/avb_paint.h"

using namespace avb::paint;

extern "C" void paint_check_render(const double* cloud, int V, const int32_t* faces, int F, const uint8_t* vpart, int W, int H,
                                   const float* intrin, float* depth, uint8_t* parts, int32_t* face_ids, int32_t* order_out) {
    std::vector<P2> proj(V);
    for (int i = 0; i < V; ++i) proj[i] = project(cloud[3 * (size_t)i], cloud[3 * (size_t)i + 1], cloud[3 * (size_t)i + 2], intrin[0], intrin[1], intrin[2], intrin[3]);
    // paint order: key descending, face index ascending (the device sorts the same 64-bit keys)
    std::vector<uint64_t> keys(F);
    for (int f = 0; f < F; ++f) {
        const int32_t* t = faces + 3 * (size_t)f;
        const float k = face_key(cloud[3 * (size_t)t[0] + 2], cloud[3 * (size_t)t[1] + 2], cloud[3 * (size_t)t[2] + 2]);
        keys[f] = order_key(k, f);
    }
    std::sort(keys.begin(), keys.end());
    std::vector<int32_t> order(F);
    for (int i = 0; i < F; ++i) order[i] = (int32_t)(keys[i] & 0xFFFFFFFFu);
    if (order_out) std::copy(order.begin(), order.end(), order_out);
    RenderView v{cloud, faces, proj.data(), vpart, W, H};
    std::vector<unsigned> wd(depth ? (size_t)W * H : 0, 0u), wp(parts ? (size_t)W * H : 0, 0u), wf(face_ids ? (size_t)W * H : 0, 0u);
    auto amax = [](unsigned* p, unsigned r) { if (r > *p) *p = r; };
    for (int i = F - 1; i >= 0; --i)   // any order: the maximum decides
        face_cover(v, order[i], (unsigned)i + 1, depth ? wd.data() : nullptr, parts ? wp.data() : nullptr, face_ids ? wf.data() : nullptr, amax);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const size_t px = (size_t)i * W + j;
            if (depth) depth[px] = resolve_depth(v, order.data(), wd[px], i, j);
            if (parts) parts[px] = resolve_parts(v, order.data(), wp[px], i, j);
            if (face_ids) face_ids[px] = resolve_faces(wf[px]);
        }
}

// renderLambert in rank form on the CPU: the same header code the device kernels call (render_vertex_lambert_kernel's
// per-vertex ordering restated with std::sort)
extern "C" void paint_check_lambert(const double* cloud, int V, const int32_t* faces, int F, int W, int H, const float* intrin,
                                    uint8_t* gray, float* vlam_out) {
    std::vector<P2> proj(V);
    for (int i = 0; i < V; ++i) proj[i] = project(cloud[3 * (size_t)i], cloud[3 * (size_t)i + 1], cloud[3 * (size_t)i + 2], intrin[0], intrin[1], intrin[2], intrin[3]);
    std::vector<uint64_t> keys(F);
    for (int f = 0; f < F; ++f) {
        const int32_t* t = faces + 3 * (size_t)f;
        keys[f] = order_key(face_key(cloud[3 * (size_t)t[0] + 2], cloud[3 * (size_t)t[1] + 2], cloud[3 * (size_t)t[2] + 2]), f);
    }
    std::sort(keys.begin(), keys.end());
    std::vector<int32_t> order(F), rank_of(F);
    for (int i = 0; i < F; ++i) {
        order[i] = (int32_t)(keys[i] & 0xFFFFFFFFu);
        rank_of[order[i]] = i;
    }
    std::vector<std::vector<int>> inc(V);
    for (int t = 0; t < F; ++t)
        for (int c = 0; c < 3; ++c) {
            auto& l = inc[faces[3 * (size_t)t + c]];
            if (l.empty() || l.back() != t) l.push_back(t);
        }
    std::vector<float> vlam(V);
    for (int v = 0; v < V; ++v) {
        std::vector<int> l = inc[v];
        std::sort(l.begin(), l.end(), [&](int a, int b) { return rank_of[a] < rank_of[b]; });
        double ns[3] = {0, 0, 0};
        for (int face : l) {
            const int32_t* t = faces + 3 * (size_t)face;
            double nn[3];
            face_unit_normal(cloud + 3 * (size_t)t[0], cloud + 3 * (size_t)t[1], cloud + 3 * (size_t)t[2], nn);
            for (int j = 0; j < 3; ++j)
                if (t[j] == v) { ns[0] = dadd(ns[0], nn[0]); ns[1] = dadd(ns[1], nn[1]); ns[2] = dadd(ns[2], nn[2]); }
        }
        vlam[v] = vertex_lambert(cloud + 3 * (size_t)v, ns);
    }
    if (vlam_out) std::copy(vlam.begin(), vlam.end(), vlam_out);
    RenderView rv{cloud, faces, proj.data(), nullptr, W, H};
    std::vector<unsigned> win((size_t)W * H, 0u);
    auto amax = [](unsigned* p, unsigned r) { if (r > *p) *p = r; };
    for (int i = F - 1; i >= 0; --i) face_cover_lambert(rv, order[i], (unsigned)i + 1, win.data(), amax);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) gray[(size_t)i * W + j] = resolve_lambert(rv, order.data(), vlam.data(), win[(size_t)i * W + j], i, j);
}
