/* oracle/ref_nanoflann_nn.cpp -- TEST INFRASTRUCTURE.
 * Thin wrapper that compiles the reference's OWN vendored nanoflann (include/nanoflann.hpp, v1.3.0,
 * header-only, std-only) where it lies under /root/reference and exposes the exact query the
 * reference issues in findNN(..., invert=true) (AvatarOptimizer.cpp:879-903): a
 * KDTreeSingleIndexAdaptor<L2_Simple, 3-D, leaf 10>, KNNResultSet(1), SearchParams(10).
 * The adaptor below is written for a plain xyz-interleaved double array (the memory layout of the
 * reference's column-major 3xN Eigen cloud, AvatarOptimizer.cpp:39-108).  Built only into
 * oracle/_ref/ (git-ignored); no reference source is copied into this repository. */
#include <nanoflann.hpp>
#include <cstdint>
#include <vector>

namespace {
struct XyzAdaptor {
    const double* pts;
    size_t n;
    inline size_t kdtree_get_point_count() const { return n; }
    inline double kdtree_get_pt(const size_t idx, int dim) const { return pts[3 * idx + dim]; }
    template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<double, XyzAdaptor>, XyzAdaptor, 3, int> Tree;
}  // namespace

extern "C" {
/* for each query (3 doubles) return the index of its exact nearest neighbour in pts[0..n) */
void ref_nanoflann_nn(const double* pts, int n, const double* queries, int nq, int32_t* out) {
    XyzAdaptor ad{pts, (size_t)n};
    Tree index(3, ad, nanoflann::KDTreeSingleIndexAdaptorParams(10));
    index.buildIndex();
    for (int i = 0; i < nq; ++i) {
        size_t idx = 0;
        double dist = 0;
        nanoflann::KNNResultSet<double> rs(1);
        rs.init(&idx, &dist);
        index.findNeighbors(rs, queries + 3 * (size_t)i, nanoflann::SearchParams(10));
        out[i] = (int32_t)idx;
    }
}
}
