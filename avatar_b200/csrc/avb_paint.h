// avb_paint.h -- the pixel coverage and the per-pixel values of the reference's triangle painters
// (internal/AvatarHelpers: paintTriangleBary AvatarHelpers.cpp:61-131, paintPartsTriangleNN :144-245,
// paintTriangleSingleColor :247-302), written so that a painter's-algorithm renderer can run in parallel:
//   * spans_*  enumerate exactly the pixels one call of the reference function writes (as inclusive runs);
//   * value_*  recompute what that call writes at one pixel (a pure function of the triangle and the pixel).
// "Last painter wins" then becomes: rank the faces in paint order, keep the highest rank per pixel (atomicMax), and
// resolve the winning face's value per pixel.
//
// Arithmetic follows the reference type by type (float in the barycentric painter, float sub-expressions widened to
// double in the other two) with explicit round-to-nearest operations on the device, so coverage and values are
// bit-identical to a non-FMA build of the reference.  Host code including this header must be compiled with
// -ffp-contract=off.  Header-only, __host__ __device__: tests/cpp/paint_check.cpp runs the same code on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define AVB_HD __host__ __device__ __forceinline__
#else
#define AVB_HD inline
#endif

namespace avb {
namespace paint {

#ifdef __CUDA_ARCH__
AVB_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
AVB_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
AVB_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
AVB_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
AVB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
AVB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
AVB_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
AVB_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
AVB_HD float fmul(float a, float b) { return a * b; }
AVB_HD float fadd(float a, float b) { return a + b; }
AVB_HD float fsub(float a, float b) { return a - b; }
AVB_HD float fdiv(float a, float b) { return a / b; }
AVB_HD double dmul(double a, double b) { return a * b; }
AVB_HD double dadd(double a, double b) { return a + b; }
AVB_HD double dsub(double a, double b) { return a - b; }
AVB_HD double ddiv(double a, double b) { return a / b; }
#endif

struct P2 { float x, y; };

AVB_HD int imax(int a, int b) { return a > b ? a : b; }
AVB_HD int imin(int a, int b) { return a < b ? a : b; }

// order of three (key, index) pairs as std::sort on std::pair<double, int> leaves them (key, then index, ascending)
AVB_HD void sort3(double k0, double k1, double k2, int o[3]) {
    o[0] = 0; o[1] = 1; o[2] = 2;
    const double k[3] = {k0, k1, k2};
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < 2 - pass; ++i) {
            const int u = o[i], v = o[i + 1];
            if (k[v] < k[u] || (k[v] == k[u] && v < u)) { o[i] = v; o[i + 1] = u; }
        }
}

// ---- paintTriangleBary (AvatarHelpers.cpp:61-131): rows in y, float edges, inclusive runs --------------------------
struct BarySetup {
    P2 a, b, c;      // vertices by ascending y; a.y floored, c.y ceiled (as the painter modifies its copies)
    int o[3];        // which input vertex became a, b, c
    bool empty;
};
AVB_HD BarySetup bary_setup(const P2 p[3]) {
    BarySetup s;
    sort3((double)p[0].y, (double)p[1].y, (double)p[2].y, s.o);
    s.a = p[s.o[0]]; s.b = p[s.o[1]]; s.c = p[s.o[2]];
    s.a.y = floorf(s.a.y);
    s.c.y = ceilf(s.c.y);
    s.empty = (s.a.y == s.c.y);
    return s;
}
template <class Emit>   // emit(row, first column, last column)
AVB_HD void spans_bary(const P2 p[3], int W, int H, Emit emit) {
    const BarySetup s = bary_setup(p);
    if (s.empty) return;
    const P2 a = s.a, b = s.b, c = s.c;
    const int minyi = imax((int)a.y, 0), maxyi = imin((int)c.y, H - 1), midyi = (int)floorf(b.y);
    if (a.y != b.y) {
        float mhi = fdiv(fsub(c.x, a.x), fsub(c.y, a.y));
        float bhi = fsub(a.x, fmul(a.y, mhi));
        float mlo = fdiv(fsub(b.x, a.x), fsub(b.y, a.y));
        float blo = fsub(a.x, fmul(a.y, mlo));
        if (b.x > c.x) { float t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = minyi; i <= imin(midyi, H - 1); ++i) {
            const int minxi = imax((int)floorf(fadd(fmul(mlo, (float)i), blo)), 0);
            const int maxxi = imin((int)ceilf(fadd(fmul(mhi, (float)i), bhi)), W - 1);
            if (minxi > maxxi) continue;
            emit(i, minxi, maxxi);
        }
    }
    if (b.y != c.y) {
        float mhi = fdiv(fsub(c.x, a.x), fsub(c.y, a.y));
        float bhi = fsub(a.x, fmul(a.y, mhi));
        float mlo = fdiv(fsub(c.x, b.x), fsub(c.y, b.y));
        float blo = fsub(b.x, fmul(b.y, mlo));
        if (b.x > a.x) { float t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = imax(midyi, 0) + (a.y != b.y ? 1 : 0); i <= maxyi; ++i) {
            const int minxi = imax((int)floorf(fadd(fmul(mlo, (float)i), blo)), 0);
            const int maxxi = imin((int)ceilf(fadd(fmul(mhi, (float)i), bhi)), W - 1);
            if (minxi > maxxi) continue;
            emit(i, minxi, maxxi);
        }
    }
}
// what the barycentric painter writes at (row i, column j); zv = the three vertex values in input order
AVB_HD float value_bary(const P2 p[3], const float zv[3], int i, int j, float maxz) {
    const BarySetup s = bary_setup(p);
    const P2 a = s.a, b = s.b, c = s.c;
    const float az = zv[s.o[0]], bz = zv[s.o[1]], cz = zv[s.o[2]];
    const float denom = fdiv(1.0f, fadd(fmul(fsub(b.x, c.x), fsub(a.y, c.y)), fmul(fsub(c.y, b.y), fsub(a.x, c.x))));
    const float w1v = fmul(fsub(b.x, c.x), fsub((float)i, c.y));
    const float w2v = fmul(fsub(c.x, a.x), fsub((float)i, c.y));
    const float jc = fsub((float)j, c.x);
    const float w1 = fmul(fadd(w1v, fmul(fsub(c.y, b.y), jc)), denom);
    const float w2 = fmul(fadd(w2v, fmul(fsub(a.y, c.y), jc)), denom);
    const float v = fadd(fadd(fmul(w1, az), fmul(w2, bz)), fmul(fsub(fsub(1.f, w1), w2), cz));
    return fminf(fmaxf(v, 0.0f), maxz);
}

// ---- paintTriangleSingleColor (AvatarHelpers.cpp:247-302): rows in y, double edges, std::fill excludes the end -----
template <class Emit>
AVB_HD void spans_single(const P2 p[3], int W, int H, Emit emit) {
    const BarySetup s = bary_setup(p);   // same vertex ordering and rounding of a.y / c.y
    if (s.empty) return;
    const P2 a = s.a, b = s.b, c = s.c;
    const int minyi = imax((int)a.y, 0), maxyi = imin((int)c.y, H - 1), midyi = (int)floorf(b.y);
    if (a.y != b.y) {
        double mhi = (double)fdiv(fsub(c.x, a.x), fsub(c.y, a.y));   // float expression stored in a double
        double bhi = dsub((double)a.x, dmul((double)a.y, mhi));
        double mlo = (double)fdiv(fsub(b.x, a.x), fsub(b.y, a.y));
        double blo = dsub((double)a.x, dmul((double)a.y, mlo));
        if (b.x > c.x) { double t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = minyi; i <= imin(midyi, H - 1); ++i) {
            const int minxi = imax((int)floor(dadd(dmul(mlo, (double)i), blo)), 0);
            const int maxxi = imin((int)ceil(dadd(dmul(mhi, (double)i), bhi)), W - 1);
            if (minxi > maxxi) continue;
            if (maxxi - 1 >= minxi) emit(i, minxi, maxxi - 1);   // std::fill(ptr + minxi, ptr + maxxi, color)
        }
    }
    if (b.y != c.y) {
        double mhi = (double)fdiv(fsub(c.x, a.x), fsub(c.y, a.y));
        double bhi = dsub((double)a.x, dmul((double)a.y, mhi));
        double mlo = (double)fdiv(fsub(c.x, b.x), fsub(c.y, b.y));
        double blo = dsub((double)b.x, dmul((double)b.y, mlo));
        if (b.x > a.x) { double t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = imax(midyi, 0) + 1; i <= maxyi; ++i) {
            const int minxi = imax((int)floor(dadd(dmul(mlo, (double)i), blo)), 0);
            const int maxxi = imin((int)ceil(dadd(dmul(mhi, (double)i), bhi)), W - 1);
            if (minxi > maxxi) continue;
            if (maxxi - 1 >= minxi) emit(i, minxi, maxxi - 1);
        }
    }
}

// ---- paintPartsTriangleNN (AvatarHelpers.cpp:144-245): columns in x, double edges, inclusive runs ------------------
struct PartsSetup {
    P2 a, b, c;      // vertices by ascending x; a.x floored, c.x ceiled
    int o[3];
    bool empty;
};
AVB_HD PartsSetup parts_setup(const P2 p[3]) {
    PartsSetup s;
    sort3((double)p[0].x, (double)p[1].x, (double)p[2].x, s.o);
    s.a = p[s.o[0]]; s.b = p[s.o[1]]; s.c = p[s.o[2]];
    s.a.x = floorf(s.a.x);
    s.c.x = ceilf(s.c.x);
    s.empty = (s.a.x == s.c.x);
    return s;
}
template <class Emit>   // emit(column, first row, last row)
AVB_HD void spans_parts(const P2 p[3], int W, int H, Emit emit) {
    const PartsSetup s = parts_setup(p);
    if (s.empty) return;
    const P2 a = s.a, b = s.b, c = s.c;
    const int minxi = imax((int)a.x, 0), maxxi = imin((int)c.x, W - 1), midxi = (int)floorf(b.x);
    if (a.x != b.x) {
        double mhi = (double)fdiv(fsub(c.y, a.y), fsub(c.x, a.x));
        double bhi = dsub((double)a.y, dmul((double)a.x, mhi));
        double mlo = (double)fdiv(fsub(b.y, a.y), fsub(b.x, a.x));
        double blo = dsub((double)a.y, dmul((double)a.x, mlo));
        if (b.y > c.y) { double t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = minxi; i <= imin(midxi, W - 1); ++i) {
            const int minyi = imax((int)floor(dadd(dmul(mlo, (double)i), blo)), 0);
            const int maxyi = imin((int)ceil(dadd(dmul(mhi, (double)i), bhi)), H - 1);
            if (minyi > maxyi) continue;
            emit(i, minyi, maxyi);
        }
    }
    if (b.x != c.x) {
        double mhi = (double)fdiv(fsub(c.y, a.y), fsub(c.x, a.x));
        double bhi = dsub((double)a.y, dmul((double)a.x, mhi));
        double mlo = (double)fdiv(fsub(c.y, b.y), fsub(c.x, b.x));
        double blo = dsub((double)b.y, dmul((double)b.x, mlo));
        if (b.y > a.y) { double t = mlo; mlo = mhi; mhi = t; t = blo; blo = bhi; bhi = t; }
        for (int i = imax(midxi, 0) + 1; i <= maxxi; ++i) {
            const int minyi = imax((int)floor(dadd(dmul(mlo, (double)i), blo)), 0);
            const int maxyi = imin((int)ceil(dadd(dmul(mhi, (double)i), bhi)), H - 1);
            if (minyi > maxyi) continue;
            emit(i, minyi, maxyi);
        }
    }
}
// which input vertex (0, 1, 2) the nearest-vertex painter picks at (row j, column i): squared distances in float,
// truncated to int, first strict minimum in the order a, b, c
AVB_HD int nearest_parts(const P2 p[3], int j, int i) {
    const PartsSetup s = parts_setup(p);
    const P2 a = s.a, b = s.b, c = s.c;
    const float fi = (float)i, fj = (float)j;
    const int dista = (int)fadd(fmul(fsub(a.x, fi), fsub(a.x, fi)), fmul(fsub(a.y, fj), fsub(a.y, fj)));
    const int distb = (int)fadd(fmul(fsub(b.x, fi), fsub(b.x, fi)), fmul(fsub(b.y, fj), fsub(b.y, fj)));
    const int distc = (int)fadd(fmul(fsub(c.x, fi), fsub(c.x, fi)), fmul(fsub(c.y, fj), fsub(c.y, fj)));
    if (dista < distb && dista < distc) return s.o[0];
    if (distb < distc) return s.o[1];
    return s.o[2];
}

// ---- AvatarRenderer pieces (AvatarRenderer.cpp:11-24, 41-70, 85-98) -------------------------------------------------
// getProjectedPoints: double arithmetic, stored as float
AVB_HD P2 project(double x, double y, double z, float fx, float cx, float fy, float cy) {
    P2 r;
    r.x = (float)dadd(ddiv(dmul(x, (double)fx), z), (double)cx);
    r.y = (float)dadd(ddiv(dmul(-y, (double)fy), z), (double)cy);
    return r;
}
// getOrderedFaces key: mean z as a float (faces are painted by decreasing key)
AVB_HD float face_key(double z0, double z1, double z2) { return (float)ddiv(dadd(dadd(z0, z1), z2), 3.0); }
// sort key of a face: ascending order of this key = decreasing float key, then increasing face index
AVB_HD uint64_t order_key(float key, int face) {
    union { float f; uint32_t u; } c;
    c.f = key;
    const uint32_t mono = (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);   // monotone in the float order
    return ((uint64_t)(~mono) << 32) | (uint32_t)face;
}
// |z| of the unit normal of triangle (a, b, c) in double; the renderers treat faces with |nz| < 0.1 as grazing
AVB_HD double face_zcross(const double a[3], const double b[3], const double c[3]) {
    const double ab[3] = {dsub(b[0], a[0]), dsub(b[1], a[1]), dsub(b[2], a[2])};
    const double ac[3] = {dsub(c[0], a[0]), dsub(c[1], a[1]), dsub(c[2], a[2])};
    const double nx = dsub(dmul(ab[1], ac[2]), dmul(ab[2], ac[1]));
    const double ny = dsub(dmul(ab[2], ac[0]), dmul(ab[0], ac[2]));
    const double nz = dsub(dmul(ab[0], ac[1]), dmul(ab[1], ac[0]));
    const double n2 = dadd(dadd(dmul(nx, nx), dmul(ny, ny)), dmul(nz, nz));
    const double zn = (n2 > 0.0) ? ddiv(nz, sqrt(n2)) : nz;   // Eigen normalized(): divides by the norm when it is positive
    return fabs(zn);
}

// unit face normal (b - a) x (c - a), Eigen normalized() (AvatarRenderer.cpp:127); every operation rounded once
AVB_HD void face_unit_normal(const double a[3], const double b[3], const double c[3], double n[3]) {
    const double ab[3] = {dsub(b[0], a[0]), dsub(b[1], a[1]), dsub(b[2], a[2])};
    const double ac[3] = {dsub(c[0], a[0]), dsub(c[1], a[1]), dsub(c[2], a[2])};
    n[0] = dsub(dmul(ab[1], ac[2]), dmul(ab[2], ac[1]));
    n[1] = dsub(dmul(ab[2], ac[0]), dmul(ab[0], ac[2]));
    n[2] = dsub(dmul(ab[0], ac[1]), dmul(ab[1], ac[0]));
    const double n2 = dadd(dadd(dmul(n[0], n[0]), dmul(n[1], n[1])), dmul(n[2], n[2]));
    if (n2 > 0.0) {
        const double r = sqrt(n2);
        n[0] = ddiv(n[0], r); n[1] = ddiv(n[1], r); n[2] = ddiv(n[2], r);
    }
}
AVB_HD void normalize3(double v[3]) {
    const double n2 = dadd(dadd(dmul(v[0], v[0]), dmul(v[1], v[1])), dmul(v[2], v[2]));
    if (n2 > 0.0) {
        const double r = sqrt(n2);
        v[0] = ddiv(v[0], r); v[1] = ddiv(v[1], r); v[2] = ddiv(v[2], r);
    }
}
// the value renderLambert paints at a vertex (AvatarRenderer.cpp:111-115, 137-163): nsum = the vertex's normal sum
AVB_HD float vertex_lambert(const double pos[3], double nsum[3]) {
    normalize3(nsum);
    if (nsum[2] > 0) { nsum[0] = -nsum[0]; nsum[1] = -nsum[1]; nsum[2] = -nsum[2]; }
    double ml[3] = {dsub(0.8, pos[0]), dsub(1.5, pos[1]), dsub(-1.2, pos[2])};
    double bl[3] = {dsub(-0.2, pos[0]), dsub(-1.5, pos[1]), dsub(0.4, pos[2])};
    normalize3(ml);
    normalize3(bl);
    const double dm = dadd(dadd(dmul(ml[0], nsum[0]), dmul(ml[1], nsum[1])), dmul(ml[2], nsum[2]));
    const double db = dadd(dadd(dmul(bl[0], nsum[0]), dmul(bl[1], nsum[1])), dmul(bl[2], nsum[2]));
    const float v = fmul((float)dadd(dmul(dm, 0.8), dmul(db, 0.2)), 255.f);
    return v > 0.f ? v : 0.f;
}

// ---- the renderer in rank form ----------------------------------------------------------------------------------------
struct RenderView {
    const double* cloud;       // [V][3] posed model
    const int32_t* faces;      // [F][3]
    const P2* proj;            // [V] projected vertices
    const uint8_t* vpart;      // [V] part of every vertex (part_map[assignedJoints[v][0]])
    int W, H;
};
AVB_HD void face_points(const RenderView& v, int face, P2 p[3], double zc_out[1]) {
    const int32_t* f = v.faces + 3 * (size_t)face;
    p[0] = v.proj[f[0]]; p[1] = v.proj[f[1]]; p[2] = v.proj[f[2]];
    zc_out[0] = face_zcross(v.cloud + 3 * (size_t)f[0], v.cloud + 3 * (size_t)f[1], v.cloud + 3 * (size_t)f[2]);
}
// coverage of one face painted with 1-based rank `rank` (its position in paint order + 1): amax(ptr, rank) keeps the
// highest rank per pixel.  Each winner image is optional.
template <class MaxOp>
AVB_HD void face_cover(const RenderView& v, int face, unsigned rank, unsigned* win_depth, unsigned* win_parts, unsigned* win_faces,
                       MaxOp amax) {
    P2 p[3];
    double zc;
    face_points(v, face, p, &zc);
    const bool grazing = zc < 0.1;   // AvatarRenderer.cpp:88-91, :186-188
    const int W = v.W, H = v.H;
    if (win_depth) {
        auto rows = [&](int i, int lo, int hi) { for (int j = lo; j <= hi; ++j) amax(win_depth + (size_t)i * W + j, rank); };
        if (grazing) spans_single(p, W, H, rows); else spans_bary(p, W, H, rows);
    }
    if (win_parts) {
        if (grazing) {
            spans_single(p, W, H, [&](int i, int lo, int hi) { for (int j = lo; j <= hi; ++j) amax(win_parts + (size_t)i * W + j, rank); });
        } else {
            spans_parts(p, W, H, [&](int col, int lo, int hi) { for (int j = lo; j <= hi; ++j) amax(win_parts + (size_t)j * W + col, rank); });
        }
    }
    if (win_faces) spans_single(p, W, H, [&](int i, int lo, int hi) { for (int j = lo; j <= hi; ++j) amax(win_faces + (size_t)i * W + j, rank); });
}
// values of the winning face at pixel (row i, column j); order[rank - 1] = model face; rank 0 = nothing painted
AVB_HD float resolve_depth(const RenderView& v, const int32_t* order, unsigned rank, int i, int j) {
    if (rank == 0) return 0.f;
    P2 p[3];
    double zc;
    const int face = order[rank - 1];
    face_points(v, face, p, &zc);
    if (zc < 0.1) return 0.f;   // grazing faces are painted with depth 0 (AvatarRenderer.cpp:89-91)
    const int32_t* f = v.faces + 3 * (size_t)face;
    const float zv[3] = {(float)v.cloud[3 * (size_t)f[0] + 2], (float)v.cloud[3 * (size_t)f[1] + 2], (float)v.cloud[3 * (size_t)f[2] + 2]};
    return value_bary(p, zv, i, j, 255.0f);
}
AVB_HD uint8_t resolve_parts(const RenderView& v, const int32_t* order, unsigned rank, int i, int j) {
    if (rank == 0) return 255;
    P2 p[3];
    double zc;
    const int face = order[rank - 1];
    face_points(v, face, p, &zc);
    if (zc < 0.1) return 255;
    const int32_t* f = v.faces + 3 * (size_t)face;
    return v.vpart[f[nearest_parts(p, i, j)]];
}
AVB_HD int32_t resolve_faces(unsigned rank) { return rank == 0 ? -1 : (int32_t)rank - 1; }

// renderLambert (AvatarRenderer.cpp:103-172): faces with |n_z| <= 1e-2 are not painted, the others with
// paintTriangleBary<uint8_t> and the three vertex values
template <class MaxOp>
AVB_HD void face_cover_lambert(const RenderView& v, int face, unsigned rank, unsigned* win, MaxOp amax) {
    P2 p[3];
    double zc;
    face_points(v, face, p, &zc);
    if (!(zc > 1e-2)) return;
    const int W = v.W, H = v.H;
    spans_bary(p, W, H, [&](int i, int lo, int hi) { for (int j = lo; j <= hi; ++j) amax(win + (size_t)i * W + j, rank); });
}
AVB_HD uint8_t resolve_lambert(const RenderView& v, const int32_t* order, const float* vlam, unsigned rank, int i, int j) {
    if (rank == 0) return 0;
    P2 p[3];
    double zc;
    const int face = order[rank - 1];
    face_points(v, face, p, &zc);
    const int32_t* f = v.faces + 3 * (size_t)face;
    const float lv[3] = {vlam[f[0]], vlam[f[1]], vlam[f[2]]};
    return (uint8_t)value_bary(p, lv, i, j, 255.0f);
}

}  // namespace paint
}  // namespace avb
