/* oracle/render_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's painter's-algorithm renderer:
 * AvatarRenderer::getProjectedPoints / getOrderedFaces / renderDepth / renderPartMask / renderFaces
 * (AvatarRenderer.cpp:11-24, 41-70, 72-98, 170-216) and the three triangle painters of internal/AvatarHelpers
 * (AvatarHelpers.cpp:61-131 paintTriangleBary, :144-245 paintPartsTriangleNN, :247-302 paintTriangleSingleColor).
 * Sequential, image at a time, faces painted far to near exactly as the reference loops do.  Compiled with
 * -ffp-contract=off: the reference is built without FMA (CMakeLists.txt:37,48: -O3 -funroll-loops, no -march).
 *
 * Deviation that cannot be avoided: the reference orders faces with std::sort, whose order among equal keys is
 * unspecified; here (and in the device renderer) equal keys keep ascending face index.
 * PARITY STATUS: the three painters are PINNED against the reference's own AvatarHelpers.cpp, compiled from
 * /root/reference into oracle/_ref/libref_painters.so with container-only OpenCV/Eigen stand-ins (oracle/shim,
 * oracle/ref_painters.cpp): tests/test_oracle.py::test_oracle_painter_equals_the_references_own_painters drives the
 * reference code face by face and gets the same images.  Projection, face order and the grazing test
 * (AvatarRenderer.cpp, which needs the full Eigen/OpenCV stack) are restated from source; the reference has no golden
 * image. */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace {

struct Pt { float x, y; };

// paintTriangleBary<T> (AvatarHelpers.cpp:61-131); T = float (renderDepth) or uint8_t (renderLambert)
template <class T>
void paint_bary(T* out, int W, int H, const Pt* proj, const int32_t* face, const float* zvec, float maxz = 255.0f) {
    std::pair<double, int> yf[3] = {{proj[face[0]].y, 0}, {proj[face[1]].y, 1}, {proj[face[2]].y, 2}};
    std::sort(yf, yf + 3);
    Pt a = proj[face[yf[0].second]], b = proj[face[yf[1].second]], c = proj[face[yf[2].second]];
    a.y = std::floor(a.y);
    c.y = std::ceil(c.y);
    if (a.y == c.y) return;
    const int minyi = std::max<int>(a.y, 0), maxyi = std::min<int>(c.y, H - 1), midyi = std::floor(b.y);
    const float az = zvec[yf[0].second], bz = zvec[yf[1].second], cz = zvec[yf[2].second];
    const float denom = 1.0f / ((b.x - c.x) * (a.y - c.y) + (c.y - b.y) * (a.x - c.x));
    auto row = [&](int i, float mlo, float blo, float mhi, float bhi) {
        const int minxi = std::max<int>(std::floor(mlo * i + blo), 0), maxxi = std::min<int>(std::ceil(mhi * i + bhi), W - 1);
        if (minxi > maxxi) return;
        const float w1v = (b.x - c.x) * (i - c.y), w2v = (c.x - a.x) * (i - c.y);
        T* ptr = out + (size_t)i * W;
        for (int j = minxi; j <= maxxi; ++j) {
            const float w1 = (w1v + (c.y - b.y) * (j - c.x)) * denom;
            const float w2 = (w2v + (a.y - c.y) * (j - c.x)) * denom;
            ptr[j] = T(std::min(std::max(w1 * az + w2 * bz + (1.f - w1 - w2) * cz, 0.0f), maxz));
        }
    };
    if (a.y != b.y) {
        float mhi = (c.x - a.x) / (c.y - a.y), bhi = a.x - a.y * mhi;
        float mlo = (b.x - a.x) / (b.y - a.y), blo = a.x - a.y * mlo;
        if (b.x > c.x) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = minyi; i <= std::min(midyi, H - 1); ++i) row(i, mlo, blo, mhi, bhi);
    }
    if (b.y != c.y) {
        float mhi = (c.x - a.x) / (c.y - a.y), bhi = a.x - a.y * mhi;
        float mlo = (c.x - b.x) / (c.y - b.y), blo = b.x - b.y * mlo;
        if (b.x > a.x) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = std::max(midyi, 0) + (a.y != b.y); i <= maxyi; ++i) row(i, mlo, blo, mhi, bhi);
    }
}

// paintTriangleSingleColor<T> (AvatarHelpers.cpp:247-302)
template <class T>
void paint_single(T* out, int W, int H, const Pt* proj, const int32_t* face, T color) {
    std::pair<double, int> yf[3] = {{proj[face[0]].y, 0}, {proj[face[1]].y, 1}, {proj[face[2]].y, 2}};
    std::sort(yf, yf + 3);
    Pt a = proj[face[yf[0].second]], b = proj[face[yf[1].second]], c = proj[face[yf[2].second]];
    a.y = std::floor(a.y);
    c.y = std::ceil(c.y);
    if (a.y == c.y) return;
    const int minyi = std::max<int>(a.y, 0), maxyi = std::min<int>(c.y, H - 1), midyi = std::floor(b.y);
    auto row = [&](int i, double mlo, double blo, double mhi, double bhi) {
        const int minxi = std::max<int>(std::floor(mlo * i + blo), 0), maxxi = std::min<int>(std::ceil(mhi * i + bhi), W - 1);
        if (minxi > maxxi) return;
        T* ptr = out + (size_t)i * W;
        std::fill(ptr + minxi, ptr + maxxi, color);   // the end is exclusive
    };
    if (a.y != b.y) {
        double mhi = (c.x - a.x) / (c.y - a.y), bhi = a.x - a.y * mhi;
        double mlo = (b.x - a.x) / (b.y - a.y), blo = a.x - a.y * mlo;
        if (b.x > c.x) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = minyi; i <= std::min(midyi, H - 1); ++i) row(i, mlo, blo, mhi, bhi);
    }
    if (b.y != c.y) {
        double mhi = (c.x - a.x) / (c.y - a.y), bhi = a.x - a.y * mhi;
        double mlo = (c.x - b.x) / (c.y - b.y), blo = b.x - b.y * mlo;
        if (b.x > a.x) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = std::max(midyi, 0) + 1; i <= maxyi; ++i) row(i, mlo, blo, mhi, bhi);
    }
}

// paintPartsTriangleNN (AvatarHelpers.cpp:144-245); vertex_part = part_map[assignedJoints[v][0].second]
void paint_parts(uint8_t* out, int W, int H, const Pt* proj, const int32_t* face, const int32_t* vertex_part) {
    std::pair<double, int> xf[3] = {{proj[face[0]].x, 0}, {proj[face[1]].x, 1}, {proj[face[2]].x, 2}};
    std::sort(xf, xf + 3);
    Pt a = proj[face[xf[0].second]], b = proj[face[xf[1].second]], c = proj[face[xf[2].second]];
    a.x = std::floor(a.x);
    c.x = std::ceil(c.x);
    if (a.x == c.x) return;
    const int pa = vertex_part[face[xf[0].second]], pb = vertex_part[face[xf[1].second]], pc = vertex_part[face[xf[2].second]];
    const int minxi = std::max<int>(a.x, 0), maxxi = std::min<int>(c.x, W - 1), midxi = std::floor(b.x);
    auto col = [&](int i, double mlo, double blo, double mhi, double bhi) {
        const int minyi = std::max<int>(std::floor(mlo * i + blo), 0), maxyi = std::min<int>(std::ceil(mhi * i + bhi), H - 1);
        if (minyi > maxyi) return;
        for (int j = minyi; j <= maxyi; ++j) {
            const int dista = (a.x - i) * (a.x - i) + (a.y - j) * (a.y - j);
            const int distb = (b.x - i) * (b.x - i) + (b.y - j) * (b.y - j);
            const int distc = (c.x - i) * (c.x - i) + (c.y - j) * (c.y - j);
            uint8_t& o = out[(size_t)j * W + i];
            if (dista < distb && dista < distc) o = (uint8_t)pa;
            else if (distb < distc) o = (uint8_t)pb;
            else o = (uint8_t)pc;
        }
    };
    if (a.x != b.x) {
        double mhi = (c.y - a.y) / (c.x - a.x), bhi = a.y - a.x * mhi;
        double mlo = (b.y - a.y) / (b.x - a.x), blo = a.y - a.x * mlo;
        if (b.y > c.y) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = minxi; i <= std::min(midxi, W - 1); ++i) col(i, mlo, blo, mhi, bhi);
    }
    if (b.x != c.x) {
        double mhi = (c.y - a.y) / (c.x - a.x), bhi = a.y - a.x * mhi;
        double mlo = (c.y - b.y) / (c.x - b.x), blo = b.y - b.x * mlo;
        if (b.y > a.y) { std::swap(mlo, mhi); std::swap(blo, bhi); }
        for (int i = std::max(midxi, 0) + 1; i <= maxxi; ++i) col(i, mlo, blo, mhi, bhi);
    }
}

// |z| of the normalised face normal (AvatarRenderer.cpp:88-91: ab.cross(ac).normalized().z())
double zcross_of(const double* cloud, const int32_t* f) {
    const double* a = cloud + 3 * (size_t)f[0];
    const double* b = cloud + 3 * (size_t)f[1];
    const double* c = cloud + 3 * (size_t)f[2];
    const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
    const double z = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    return std::fabs(z > 0.0 ? n[2] / std::sqrt(z) : n[2]);
}

}  // namespace

extern "C" {

/* cloud [V][3] fp64 (the posed model, Avatar::cloud), faces [F][3], vertex_part [V], intrin = {fx, cx, fy, cy}.
 * Outputs (each nullable): depth float [H][W] (0 = nothing), parts uint8 [H][W] (255 = nothing), face_ids int32 [H][W]
 * (-1 = nothing; as in the reference the value is the face's position in paint order, order_out[value] is the model
 * face).  order_out (nullable, F entries): model faces in paint order. */
void orc_render(const double* cloud, int V, const int32_t* faces, int F, const int32_t* vertex_part, int W, int H,
                const float* intrin, float* depth, uint8_t* parts, int32_t* face_ids, int32_t* order_out) {
    const float fx = intrin[0], cx = intrin[1], fy = intrin[2], cy = intrin[3];
    std::vector<Pt> proj(V);   // getProjectedPoints (AvatarRenderer.cpp:11-24)
    for (int i = 0; i < V; ++i) {
        const double* pt = cloud + 3 * (size_t)i;
        proj[i].x = static_cast<double>(pt[0]) * fx / pt[2] + cx;
        proj[i].y = -static_cast<double>(pt[1]) * fy / pt[2] + cy;
    }
    // getOrderedFaces (AvatarRenderer.cpp:41-70): by decreasing mean z (float key); equal keys keep face order
    std::vector<std::pair<float, int>> ord(F);
    for (int i = 0; i < F; ++i) {
        const int32_t* f = faces + 3 * (size_t)i;
        ord[i].first = (cloud[3 * (size_t)f[0] + 2] + cloud[3 * (size_t)f[1] + 2] + cloud[3 * (size_t)f[2] + 2]) / 3.f;
        ord[i].second = i;
    }
    std::stable_sort(ord.begin(), ord.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
    if (order_out)
        for (int i = 0; i < F; ++i) order_out[i] = ord[i].second;
    if (depth) std::fill(depth, depth + (size_t)W * H, 0.f);          // cv::Mat::zeros (AvatarRenderer.cpp:82)
    if (parts) std::fill(parts, parts + (size_t)W * H, (uint8_t)255);  // setTo(255) (:180)
    if (face_ids) std::fill(face_ids, face_ids + (size_t)W * H, -1);   // setTo(-1) (:207)
    for (int i = 0; i < F; ++i) {
        const int32_t* f = faces + 3 * (size_t)ord[i].second;
        const double zc = zcross_of(cloud, f);
        if (depth) {   // renderDepth (AvatarRenderer.cpp:72-98)
            if (zc < 0.1) {
                paint_single<float>(depth, W, H, proj.data(), f, 0.f);
            } else {
                const float zv[3] = {(float)cloud[3 * (size_t)f[0] + 2], (float)cloud[3 * (size_t)f[1] + 2], (float)cloud[3 * (size_t)f[2] + 2]};
                paint_bary<float>(depth, W, H, proj.data(), f, zv);
            }
        }
        if (parts) {   // renderPartMask (AvatarRenderer.cpp:170-197)
            if (zc < 0.1) paint_single<uint8_t>(parts, W, H, proj.data(), f, (uint8_t)255);
            else paint_parts(parts, W, H, proj.data(), f, vertex_part);
        }
        if (face_ids) paint_single<int32_t>(face_ids, W, H, proj.data(), f, i);   // renderFaces (:199-216) paints the position in paint order
    }
}


/* AvatarRenderer::renderLambert (AvatarRenderer.cpp:103-172): per-vertex normals = sum of the unit normals of the incident
 * faces, added IN PAINT ORDER (the loop runs over the ordered faces), normalised, flipped towards the camera (z <= 0);
 * two point lights; faces with |n_z| <= 1e-2 are skipped; painted far to near with paintTriangleBary<uint8_t>.
 * Eigen conventions restated: cross = (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0); normalized() / normalize() divide by
 * sqrt(squaredNorm) when squaredNorm > 0 and leave the vector otherwise; dot and squaredNorm sum ((x + y) + z); abs() is
 * the floating overload.  gray [H][W] uint8 (0 = nothing).  Optional taps: vertex_lambert [V] float (the value painted at
 * every vertex), face_visible [F] (ordered faces). */
void orc_render_lambert(const double* cloud, int V, const int32_t* faces, int F, int W, int H, const float* intrin, uint8_t* gray,
                        float* vertex_lambert, uint8_t* face_visible) {
    const float fx = intrin[0], cx = intrin[1], fy = intrin[2], cy = intrin[3];
    std::vector<Pt> proj(V);
    for (int i = 0; i < V; ++i) {
        const double* pt = cloud + 3 * (size_t)i;
        proj[i].x = static_cast<double>(pt[0]) * fx / pt[2] + cx;
        proj[i].y = -static_cast<double>(pt[1]) * fy / pt[2] + cy;
    }
    std::vector<std::pair<float, int>> ord(F);
    for (int i = 0; i < F; ++i) {
        const int32_t* f = faces + 3 * (size_t)i;
        ord[i].first = (cloud[3 * (size_t)f[0] + 2] + cloud[3 * (size_t)f[1] + 2] + cloud[3 * (size_t)f[2] + 2]) / 3.f;
        ord[i].second = i;
    }
    std::stable_sort(ord.begin(), ord.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
    auto normalize3 = [](double* v) {
        const double z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
        if (z > 0.0) {
            const double n = std::sqrt(z);
            v[0] /= n; v[1] /= n; v[2] /= n;
        }
    };
    std::vector<double> vn(3 * (size_t)V, 0.0);
    std::vector<uint8_t> visible(F);
    for (int i = 0; i < F; ++i) {
        const int32_t* f = faces + 3 * (size_t)ord[i].second;
        const double* a = cloud + 3 * (size_t)f[0];
        const double* b = cloud + 3 * (size_t)f[1];
        const double* c = cloud + 3 * (size_t)f[2];
        const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
        normalize3(n);
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) vn[3 * (size_t)f[j] + k] += n[k];
        visible[i] = std::fabs(n[2]) > 1e-2;
    }
    const double mainLight[3] = {0.8, 1.5, -1.2}, backLight[3] = {-0.2, -1.5, 0.4};
    const double mainI = 0.8, backI = 0.2;
    std::vector<float> lam(V);
    for (int v = 0; v < V; ++v) {
        double* n = &vn[3 * (size_t)v];
        normalize3(n);
        if (n[2] > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
        const double* a = cloud + 3 * (size_t)v;
        double ml[3] = {mainLight[0] - a[0], mainLight[1] - a[1], mainLight[2] - a[2]};
        double bl[3] = {backLight[0] - a[0], backLight[1] - a[1], backLight[2] - a[2]};
        normalize3(ml);
        normalize3(bl);
        const double dm = (ml[0] * n[0] + ml[1] * n[1]) + ml[2] * n[2], db = (bl[0] * n[0] + bl[1] * n[1]) + bl[2] * n[2];
        lam[v] = std::max(float(dm * mainI + db * backI) * 255, 0.f);
    }
    if (vertex_lambert) std::copy(lam.begin(), lam.end(), vertex_lambert);
    if (face_visible) std::copy(visible.begin(), visible.end(), face_visible);
    std::fill(gray, gray + (size_t)W * H, (uint8_t)0);
    for (int i = 0; i < F; ++i) {
        if (!visible[i]) continue;
        const int32_t* f = faces + 3 * (size_t)ord[i].second;
        const float lv[3] = {lam[f[0]], lam[f[1]], lam[f[2]]};
        paint_bary<uint8_t>(gray, W, H, proj.data(), f, lv);
    }
}

/* the renderer's per-frame prelude, for tests that drive the reference's own painters (oracle/_ref/libref_painters.so):
 * projected vertices [V][2], faces in paint order [F], and per ordered face the grazing flag (|n_z| < 0.1) */
void orc_render_prelude(const double* cloud, int V, const int32_t* faces, int F, const float* intrin, float* proj_xy,
                        int32_t* order, uint8_t* grazing) {
    const float fx = intrin[0], cx = intrin[1], fy = intrin[2], cy = intrin[3];
    for (int i = 0; i < V; ++i) {
        const double* pt = cloud + 3 * (size_t)i;
        proj_xy[2 * i] = static_cast<double>(pt[0]) * fx / pt[2] + cx;
        proj_xy[2 * i + 1] = -static_cast<double>(pt[1]) * fy / pt[2] + cy;
    }
    std::vector<std::pair<float, int>> ord(F);
    for (int i = 0; i < F; ++i) {
        const int32_t* f = faces + 3 * (size_t)i;
        ord[i].first = (cloud[3 * (size_t)f[0] + 2] + cloud[3 * (size_t)f[1] + 2] + cloud[3 * (size_t)f[2] + 2]) / 3.f;
        ord[i].second = i;
    }
    std::stable_sort(ord.begin(), ord.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
    for (int i = 0; i < F; ++i) {
        order[i] = ord[i].second;
        grazing[i] = zcross_of(cloud, faces + 3 * (size_t)ord[i].second) < 0.1 ? 1 : 0;
    }
}
}  // extern "C"
