// include/ark/Util.h -- the two helpers of the reference's include/Util.h that its callers' fit sections use
#pragma once
#include "../ark_b200/AvatarOptimizer.h"
namespace ark {
namespace util {
/** Util.cpp paletteColor: a fixed colour per body part (visualisation only; the exact palette is not reproduced) */
inline cv::Vec3b paletteColor(int color_index, bool bgr = true) {
    static const uint8_t pal[8][3] = {{255, 87, 51}, {51, 255, 87}, {51, 87, 255}, {255, 215, 0}, {255, 0, 255}, {0, 255, 255}, {160, 82, 45}, {128, 128, 128}};
    const uint8_t* c = pal[((color_index % 8) + 8) % 8];
    return bgr ? cv::Vec3b(c[2], c[1], c[0]) : cv::Vec3b(c[0], c[1], c[2]);
}
}  // namespace util
}  // namespace ark
