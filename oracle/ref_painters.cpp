/* oracle/ref_painters.cpp -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE'S OWN triangle painters:
 * this file is compiled together with /root/reference/AvatarHelpers.cpp (the source where it lies, never copied)
 * against the container-only stand-ins of oracle/shim, into oracle/_ref/libref_painters.so.  It pins
 * oracle/render_oracle.cpp (and through it the device renderer) against the reference's real scan-line code. */
#include <cstdint>
#include <utility>
#include <vector>
#include "Calibration.h"
#include "internal/AvatarHelpers.h"

namespace {
std::vector<cv::Point2f> points_of(const float* xy, int n) {
    std::vector<cv::Point2f> p((size_t)n);
    for (int i = 0; i < n; ++i) p[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
    return p;
}
}  // namespace

extern "C" {

/* paint `nf` faces (rows of faces[nf][3]) in the given order into the images, as AvatarRenderer::renderDepth /
 * renderPartMask / renderFaces do with the reference painters: grazing[i] != 0 selects the single-colour painter
 * (depth 0 / part 255); face_ids gets the loop index.  proj: [V][2] projected vertices; vertex_part: [V]. */
void ref_paint_faces(const float* proj_xy, int V, const int32_t* faces, int nf, const uint8_t* grazing, const float* zv /*[nf][3]*/,
                     const int32_t* vertex_part, int W, int H, float* depth, uint8_t* parts, int32_t* face_ids) {
    const std::vector<cv::Point2f> proj = points_of(proj_xy, V);
    const cv::Size size(W, H);
    std::vector<std::vector<std::pair<double, int>>> assigned((size_t)V);
    for (int v = 0; v < V; ++v) assigned[v].push_back({1.0, vertex_part[v]});   // [0].second is all the painter reads
    const std::vector<int> no_part_map;                                          // empty: joints are already parts
    cv::Mat md(H, W, sizeof(float), depth), mp(H, W, 1, parts), mf(H, W, sizeof(int32_t), face_ids);
    for (int i = 0; i < nf; ++i) {
        const cv::Vec3i face(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]);
        if (depth) {
            if (grazing[i]) ark::paintTriangleSingleColor<float>(md, size, proj, face, 0.f);
            else ark::paintTriangleBary<float>(md, size, proj, face, zv + 3 * i);
        }
        if (parts) {
            if (grazing[i]) ark::paintTriangleSingleColor<uint8_t>(mp, size, proj, face, uint8_t(255));
            else ark::paintPartsTriangleNN(mp, size, proj, assigned, face, no_part_map);
        }
        if (face_ids) ark::paintTriangleSingleColor<int>(mf, size, proj, face, i);
    }
}


/* renderLambert's painting loop with the reference's paintTriangleBary<uint8_t>: faces in the given (paint) order, only
 * those with visible[i] != 0, lambert[i][3] = the three vertex values of face i */
void ref_paint_lambert(const float* proj_xy, int V, const int32_t* faces, int nf, const uint8_t* visible, const float* lambert,
                       int W, int H, uint8_t* gray) {
    const std::vector<cv::Point2f> proj = points_of(proj_xy, V);
    const cv::Size size(W, H);
    cv::Mat mg(H, W, 1, gray);
    for (int i = 0; i < nf; ++i) {
        if (!visible[i]) continue;
        const cv::Vec3i face(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]);
        ark::paintTriangleBary<uint8_t>(mg, size, proj, face, lambert + 3 * i);
    }
}

/* CameraIntrin::depthToXYZ of the reference (Calibration.cpp:83-95): depth [H][W] float -> xyz [H][W][3] float */
void ref_depth_to_xyz(const float* depth, int W, int H, float fx, float cx, float fy, float cy, float* xyz) {
    ark::CameraIntrin k;
    k.clear();
    k.fx = fx; k.cx = cx; k.fy = fy; k.cy = cy;
    const cv::Mat d(H, W, sizeof(float), const_cast<float*>(depth));
    const cv::Mat out = k.depthToXYZ(d);
    std::memcpy(xyz, out.data, (size_t)W * H * 12);
}

/* CameraIntrin::readFile / writeFile of the reference (Calibration.cpp:19-51, 97-111): read path_in, optionally write
 * path_out, return fx, fy, cx, cy, k[6], p[2]; result 0 iff readFile returned true */
int ref_intrin_probe(const char* path_in, const char* path_out, float* out12) {
    ark::CameraIntrin c;
    const bool ok = c.readFile(path_in);
    out12[0] = c.fx; out12[1] = c.fy; out12[2] = c.cx; out12[3] = c.cy;
    for (int i = 0; i < 6; ++i) out12[4 + i] = c.k[i];
    out12[10] = c.p[0]; out12[11] = c.p[1];
    if (path_out && !c.writeFile(path_out)) return 2;
    return ok ? 0 : 1;
}
}  // extern "C"
