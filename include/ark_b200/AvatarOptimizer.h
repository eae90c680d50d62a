// ark_b200/AvatarOptimizer.h -- header-compatible stand-in for include/AvatarOptimizer.h:11-54.
#pragma once
#include "Avatar.h"
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#endif

namespace ark {
#if !__has_include(<opencv2/core.hpp>)
/** include/Calibration.h:11-76: pinhole intrinsics with the reference's text file format (Calibration.cpp:19-51,
 *  97-111; tags fx/fy/cx/cy, k1..k6, p1..p2 on reading -- the writer emits 0-based k/p tags, a reference quirk kept) */
struct CameraIntrin {
    float fx = 0, fy = 0, cx = 0, cy = 0;
    float k[6] = {0, 0, 0, 0, 0, 0};
    float p[2] = {0, 0};
    CameraIntrin() {}
    explicit CameraIntrin(const std::string& path) { readFile(path); }
    void clear();
    bool readFile(const std::string& path);          // true iff fx, fy, cx, cy were all present
    bool writeFile(const std::string& path) const;
    /** Calibration.cpp:68-74 / :76-80, float arithmetic */
    void to3D(float px, float py, float depth, float out_xyz[3]) const;
    void to2D(const float xyz[3], float out_xy[2]) const;
};
#endif

class AvatarOptimizer {
   public:
    AvatarOptimizer(Avatar& ava, const CameraIntrin& intrin, const cv::Size& image_size, int num_parts,
                    const std::vector<int>& part_map);
    ~AvatarOptimizer();
    void optimize(const Eigen::Matrix<double, 3, Eigen::Dynamic>& data_cloud, const Eigen::VectorXi& data_part_labels,
                  int icp_iters = 1, int num_threads = 4);
    static const int ROT_SIZE = 4;
    std::vector<Eigen::Quaterniond, Eigen::aligned_allocator<Eigen::Quaterniond>> r;
    double betaPose = 0.1, betaShape = 1.0;
    int nnStep = 20;
    int maxItersPerICP = 10;
    bool enableOcclusion = true;
    Avatar& ava;
    const CameraIntrin& intrin;
    cv::Size imageSize;
    int numParts;
    const std::vector<int>& partMap;
    /** extension: statistics of the last optimize() call */
    int lastIterations = 0, lastCorrespondences = 0;
    double lastInitialCost = 0, lastFinalCost = 0;
   private:
    avb_fitter* fitter_ = nullptr;
    int capacity_ = 0;
};
}  // namespace ark
