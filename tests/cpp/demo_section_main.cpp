// demo_section_main.cpp -- drives the fit section of the reference's own demo.cpp (tests/cpp/_gen/demo_fit_section.cpp, generated
// from /root/reference at test time) compiled against the facade: model dir + RTree file + one depth frame in, fitted
// parameters and a checksum of the Lambert overlay out.  Used by tests/test_cpp_facade.py.
#include "Avatar.h"
#include "AvatarOptimizer.h"
#include "AvatarRenderer.h"
#include "RTree.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct BGSubtractorView { cv::Point topLeft, botRight; };
cv::Mat demo_fit_section(ark::RTree& rtree, ark::Avatar& ava, ark::AvatarOptimizer& avaOpt, const ark::CameraIntrin& intrin,
                         cv::Mat& depth, cv::Mat& image, cv::Mat& vis, BGSubtractorView& bgsub,
                         Eigen::Matrix<double, 2, Eigen::Dynamic>& comPre, int interval, int frameICPIters, int reinitICPIters,
                         size_t reinitCnz, bool& reinit, bool rtreeOnly, cv::Mat* labels_out);

int main(int argc, char** argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: demo_section <model_dir> <tree.srtr> <frame.bin> <frames>\n");
        return 2;
    }
    ark::RTree rtree{std::string(argv[2])};
    ark::AvatarModel avaModel(argv[1]);
    ark::Avatar ava(avaModel);
    FILE* fp = std::fopen(argv[3], "rb");
    if (!fp) return 3;
    int32_t hdr[8];   // W, H, x0, y0, x1, y1, interval, icp iterations
    float k[4];       // fx, cx, fy, cy
    if (std::fread(hdr, 4, 8, fp) != 8 || std::fread(k, 4, 4, fp) != 4) return 3;
    const int W = hdr[0], H = hdr[1];
    cv::Mat depth(H, W, CV_32F), image(H, W, CV_32FC3), vis(H, W, CV_8UC3);
    if (std::fread(depth.data, 4, (size_t)W * H, fp) != (size_t)W * H) return 3;
    std::fclose(fp);
    ark::CameraIntrin intrin;
    intrin.fx = k[0]; intrin.cx = k[1]; intrin.fy = k[2]; intrin.cy = k[3];
    for (int r = 0; r < H; ++r)   // the xyz map the camera delivers (CameraIntrin::depthToXYZ, Calibration.cpp:83-95)
        for (int c = 0; c < W; ++c) {
            float xyz[3];
            intrin.to3D((float)c, (float)r, depth.at<float>(r, c), xyz);
            image.at<cv::Vec3f>(r, c) = cv::Vec3f(xyz[0], xyz[1], xyz[2]);
        }
    ark::AvatarOptimizer avaOpt(ava, intrin, cv::Size(W, H), rtree.numParts, rtree.partMap);
    BGSubtractorView bgsub{cv::Point(hdr[2], hdr[3]), cv::Point(hdr[4], hdr[5])};
    Eigen::Matrix<double, 2, Eigen::Dynamic> comPre;
    bool reinit = true;
    const int frames = std::atoi(argv[4]);
    for (int t = 0; t < frames; ++t) {
        demo_fit_section(rtree, ava, avaOpt, intrin, depth, image, vis, bgsub, comPre, hdr[6], hdr[7], hdr[7] + 2, 1000, reinit, false, nullptr);
        const std::vector<double> x = ava.packParams();
        std::printf("PARAMS");
        for (double v : x) std::printf(" %.17g", v);
        unsigned long long sum = 0, lit = 0;
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) {
                sum += (unsigned long long)vis.at<cv::Vec3b>(r, c)[0] * (unsigned)(1 + (r * 31 + c) % 97);
                lit += vis.at<cv::Vec3b>(r, c)[0] > 0;
            }
        std::printf("\nOVERLAY %llu %llu\nCOM", sum, lit);
        for (int i = 0; i < (int)comPre.cols(); ++i) std::printf(" %.17g %.17g", comPre(0, i), comPre(1, i));
        std::printf("\n");
    }
    return 0;
}
