// facade_demo.cpp -- a reference-style caller (demo.cpp:135-143, 254-268) written against the facade headers:
// model directory (model.npz + pose_prior.txt) -> Avatar -> AvatarOptimizer::optimize on a cloud read from a
// binary frame file; prints the fitted parameters.  Used by tests/test_cpp_facade.py.
#include <ark_b200/AvatarOptimizer.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: facade_demo <model_dir> <frame.bin> <icp_iters> [partmap16...]\n");
        return 2;
    }
    using namespace ark;
    const AvatarModel model(argv[1]);
    if (std::string(argv[2]) == "--info") {  // host-only: what the npz / pose_prior.txt readers produced
        double s0 = 0, s1 = 0, s2 = 0;
        for (long i = 0; i < model.baseCloud.rows(); ++i) s0 += model.baseCloud(i);
        for (long i = 0; i < model.jointShapeReg.rows(); ++i)
            for (long k = 0; k < model.jointShapeReg.cols(); ++k) s1 += model.jointShapeReg(i, k);
        for (auto& a : model.assignedJoints) s2 += a[0].first + a[0].second;
        std::printf("INFO %d %d %d %d %d %d %.17g %.17g %.17g %.17g\n", model.numPoints(), model.numJoints(),
                    model.numShapeKeys(), model.numFaces(), model.posePrior.nComps, model.posePrior.nDims, s0, s1, s2,
                    model.hasPosePrior() ? model.posePrior.cov[1](3, 4) : 0.0);
        return 0;
    }
    Avatar ava(model);
    if (std::string(argv[2]) == "--align") {  // host-only: Avatar::alignToJoints + smplParams on 24 joint positions from a file
        FILE* fj = std::fopen(argv[3], "rb");
        if (!fj) return 3;
        std::vector<double> jp(3 * 24);
        if (std::fread(jp.data(), 8, jp.size(), fj) != jp.size()) return 3;
        std::fclose(fj);
        CloudType pos(3, 24);
        for (int j = 0; j < 24; ++j)
            for (int c = 0; c < 3; ++c) pos(c, j) = jp[3 * (size_t)j + c];
        ava.alignToJoints(pos);
        std::printf("P %.17g %.17g %.17g\nW0 %.17g\nR", ava.p(0), ava.p(1), ava.p(2), ava.w(0));
        for (int j = 0; j < 24; ++j)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) std::printf(" %.17g", ava.r[(size_t)j](a, b));
        const Eigen::VectorXd sp = ava.smplParams();
        std::printf("\nSMPL");
        for (long i = 0; i < sp.rows(); ++i) std::printf(" %.17g", sp(i));
        std::printf("\n");
        return 0;
    }
    FILE* fp = std::fopen(argv[2], "rb");
    if (!fp) return 3;
    int32_t hdr[4];  // N, J, K, numParts
    if (std::fread(hdr, 4, 4, fp) != 4) return 3;
    const int N = hdr[0], J = hdr[1], K = hdr[2], numParts = hdr[3];
    std::vector<int> partMap(J);
    std::vector<double> x(3 + 4 * J + K), pts(3 * (size_t)N);
    std::vector<int32_t> lab(N);
    if (std::fread(partMap.data(), 4, J, fp) != (size_t)J || std::fread(x.data(), 8, x.size(), fp) != x.size() ||
        std::fread(pts.data(), 8, pts.size(), fp) != pts.size() || std::fread(lab.data(), 4, N, fp) != (size_t)N)
        return 3;
    std::fclose(fp);
    ava.unpackParams(x);   // p, r (rotation matrices), w
    ava.update();
    CameraIntrin intrin;
    AvatarOptimizer avaOpt(ava, intrin, cv::Size(640, 576), numParts, partMap);
    avaOpt.betaPose = 0.05;   // demo.cpp:54-57
    avaOpt.betaShape = 0.12;
    Eigen::Matrix<double, 3, Eigen::Dynamic> dataCloud(3, N);
    Eigen::VectorXi dataPartLabels(N);
    for (int i = 0; i < N; ++i) {
        for (int c = 0; c < 3; ++c) dataCloud(c, i) = pts[3 * (size_t)i + c];
        dataPartLabels(i) = lab[i];
    }
    avaOpt.optimize(dataCloud, dataPartLabels, std::atoi(argv[3]), 4);
    const std::vector<double> out = ava.packParams();
    std::printf("PARAMS");
    for (double v : out) std::printf(" %.17g", v);
    std::printf("\nCLOUD0 %.17g %.17g %.17g\n", ava.cloud(0, 0), ava.cloud(1, 0), ava.cloud(2, 0));
    std::printf("STATS %d %d %.17g %.17g\n", avaOpt.lastIterations, avaOpt.lastCorrespondences, avaOpt.lastInitialCost,
                avaOpt.lastFinalCost);
    return 0;
}
