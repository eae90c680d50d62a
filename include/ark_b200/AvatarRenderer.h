// ark_b200/AvatarRenderer.h -- header-compatible stand-in for include/AvatarRenderer.h:13-70 of the reference, backed by
// avb_render_batch / avb_render_lambert_batch: the same constructor and render* members, cv::Mat results.
#pragma once
#include "AvatarOptimizer.h"

namespace ark {
class AvatarRenderer {
   public:
    AvatarRenderer(const Avatar& ava, const CameraIntrin& intrin);
    ~AvatarRenderer();
    AvatarRenderer(const AvatarRenderer&) = delete;
    /** AvatarRenderer.cpp:72-98: CV_32F, 0 = nothing */
    cv::Mat renderDepth(const cv::Size& image_size) const;
    /** AvatarRenderer.cpp:103-172: CV_8U Lambertian shading, 0 = nothing */
    cv::Mat renderLambert(const cv::Size& image_size) const;
    /** AvatarRenderer.cpp:174-197: CV_8U part of the nearest projected vertex, 255 = nothing */
    cv::Mat renderPartMask(const cv::Size& image_size, const std::vector<int>& part_map) const;
    /** AvatarRenderer.cpp:199-216: CV_32S position of the face in paint order, -1 = nothing */
    cv::Mat renderFaces(const cv::Size& image_size, int num_threads = 1) const;
    /** AvatarRenderer.cpp:218-222: the reference caches projections and the face order; nothing is cached here */
    void update() const {}
    const Avatar& ava;
    const CameraIntrin& intrin;
   private:
    avb_fitter* fitter(const std::vector<int>& part_map) const;
    mutable avb_fitter* fitter_ = nullptr;
    mutable std::vector<int> fitter_part_map_;
};
}  // namespace ark
