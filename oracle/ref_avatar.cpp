// oracle/ref_avatar.cpp -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own Avatar.cpp and GaussianMixture.cpp,
// which oracle/Makefile compiles from /root/reference (never copied) against the Eigen stand-in under oracle/shim.  The
// tests use it to check that the oracle's restatement of Avatar::update (Avatar.cpp:22-75), GaussianMixture::load / residual /
// pdf (GaussianMixture.cpp:12-114), Avatar::smplParams (:128-137) and Avatar::alignToJoints (:141-193) computes what the
// reference's code computes.  AvatarModel's file loader (AvatarModel.cpp: cnpy, PCD readers) is not compiled; its
// constructor is defined here as "empty model" and the fields are filled from arrays.
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "Avatar.h"
#include "GaussianMixture.h"
#include "Util.h"

namespace ark {
AvatarModel::AvatarModel(const std::string& model_dir, bool) : MODEL_DIR(model_dir) {
    useJointShapeRegressor = false;
    posePrior.nComps = -1;
}
namespace random_util {   // Util.cpp is not compiled; deterministic stand-ins (the tests do not sample)
static std::mt19937& gen() { static std::mt19937 g(12345); return g; }
float uniform(float a, float b) { return std::uniform_real_distribution<float>(a, b)(gen()); }
float randn(float m, float v) { return std::normal_distribution<float>(m, v)(gen()); }
float uniform(std::mt19937& rg, float a, float b) { return std::uniform_real_distribution<float>(a, b)(rg); }
float randn(std::mt19937& rg, float m, float v) { return std::normal_distribution<float>(m, v)(rg); }
}  // namespace random_util
}  // namespace ark

namespace {
struct RefModel {
    ark::AvatarModel m;
    RefModel() : m("") {}
};
}  // namespace

extern "C" {

// ---- GaussianMixture ----
void* ref_gmm_load(const char* path) {
    auto* g = new ark::GaussianMixture;
    g->load(path);
    return g;
}
void ref_gmm_free(void* h) { delete static_cast<ark::GaussianMixture*>(h); }
int ref_gmm_dims(void* h, int* ncomps, int* ndims) {
    auto* g = static_cast<ark::GaussianMixture*>(h);
    *ncomps = g->nComps;
    *ndims = g->nDims;
    return 0;
}
// residual [nDims + 1], returns the component index
int ref_gmm_residual(void* h, const double* x, double* out) {
    auto* g = static_cast<ark::GaussianMixture*>(h);
    Eigen::VectorXd xv(g->nDims);
    for (int i = 0; i < g->nDims; ++i) xv[i] = x[i];
    int comp = -1;
    Eigen::VectorXd r = g->residual(xv, &comp);
    for (int i = 0; i <= g->nDims; ++i) out[i] = r[i];
    return comp;
}
double ref_gmm_pdf(void* h, const double* x) {
    auto* g = static_cast<ark::GaussianMixture*>(h);
    Eigen::VectorXd xv(g->nDims);
    for (int i = 0; i < g->nDims; ++i) xv[i] = x[i];
    return g->pdf(xv);
}
// prec_cho [C][D][D] row-major, consts_log [C]
void ref_gmm_tables(void* h, double* prec_cho, double* consts_log) {
    auto* g = static_cast<ark::GaussianMixture*>(h);
    const int D = g->nDims;
    for (int c = 0; c < g->nComps; ++c) {
        consts_log[c] = g->consts_log[c];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) prec_cho[((size_t)c * D + i) * D + j] = g->prec_cho[(size_t)c](i, j);
    }
}

// ---- AvatarModel from arrays: v_template [V][3], shapedirs [V][3][K], j_regressor [J][V], weights [V][J], parent [J] ----
void* ref_model_create(int V, int J, int K, const double* v_template, const double* shapedirs, const double* j_regressor,
                       const double* weights, const int32_t* parent) {
    auto* rm = new RefModel;
    ark::AvatarModel& m = rm->m;
    m.baseCloud.resize(3 * V);
    for (int i = 0; i < 3 * V; ++i) m.baseCloud[i] = v_template[i];
    m.keyClouds.resize(3 * V, K);
    for (int i = 0; i < 3 * V; ++i)
        for (int k = 0; k < K; ++k) m.keyClouds(i, k) = shapedirs[(size_t)i * K + k];
    m.parent.resize(J);
    for (int j = 0; j < J; ++j) m.parent[j] = parent[j];
    std::vector<Eigen::Triplet<double>> tj, tw;
    for (int j = 0; j < J; ++j)
        for (int v = 0; v < V; ++v)
            if (j_regressor[(size_t)j * V + v] != 0.0) tj.emplace_back(v, j, j_regressor[(size_t)j * V + v]);
    m.jointRegressor.resize(V, J);
    m.jointRegressor.setFromTriplets(tj.begin(), tj.end());
    for (int v = 0; v < V; ++v)
        for (int j = 0; j < J; ++j)
            if (weights[(size_t)v * J + j] != 0.0) tw.emplace_back(j, v, weights[(size_t)v * J + j]);
    m.weights.resize(J, V);
    m.weights.setFromTriplets(tw.begin(), tw.end());
    // AvatarModel.cpp: initialJointPos = baseCloud (as 3 x V) * jointRegressor
    Eigen::Map<ark::CloudType> base(m.baseCloud.data(), 3, V);
    m.initialJointPos = base * m.jointRegressor;
    return rm;
}
void ref_model_free(void* h) { delete static_cast<RefModel*>(h); }

// Avatar::update with p [3], r [J][9] row-major rotation matrices, w [K] -> cloud [V][3], jointPos [J][3], jointTrans [J][12]
// (column-major 3 x 4 per joint, as the reference stores it)
void ref_avatar_update(void* h, const double* p, const double* r, const double* w, double* cloud, double* joint_pos,
                       double* joint_trans) {
    ark::AvatarModel& m = static_cast<RefModel*>(h)->m;
    ark::Avatar ava(m);
    const int J = m.numJoints(), V = m.numPoints(), K = m.numShapeKeys();
    for (int i = 0; i < 3; ++i) ava.p[i] = p[i];
    for (int k = 0; k < K; ++k) ava.w[k] = w[k];
    for (int j = 0; j < J; ++j)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) ava.r[(size_t)j](a, b) = r[(size_t)j * 9 + 3 * a + b];
    FILE* keep = stdout;   // update() prints its profile line on every call
    stdout = fopen("/dev/null", "w");
    ava.update();
    fclose(stdout);
    stdout = keep;
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < 3; ++c) cloud[(size_t)v * 3 + c] = ava.cloud(c, v);
    for (int j = 0; j < J; ++j) {
        for (int c = 0; c < 3; ++c) joint_pos[(size_t)j * 3 + c] = ava.jointPos(c, j);
        for (int e = 0; e < 12; ++e) joint_trans[(size_t)j * 12 + e] = ava.jointTrans(e, j);
    }
}

// Avatar::alignToJoints(pos [24][3]) then smplParams(): p [3], r [J][9], w0, smpl [3 (J - 1)]
void ref_avatar_align(void* h, const double* pos, double* p, double* r, double* w0, double* smpl) {
    ark::AvatarModel& m = static_cast<RefModel*>(h)->m;
    ark::Avatar ava(m);
    const int J = m.numJoints();
    ark::CloudType P(3, J);
    for (int j = 0; j < J; ++j)
        for (int c = 0; c < 3; ++c) P(c, j) = pos[(size_t)j * 3 + c];
    ava.alignToJoints(P);
    for (int i = 0; i < 3; ++i) p[i] = ava.p[i];
    for (int j = 0; j < J; ++j)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) r[(size_t)j * 9 + 3 * a + b] = ava.r[(size_t)j](a, b);
    *w0 = ava.w[0];
    Eigen::VectorXd s = ava.smplParams();
    for (int i = 0; i < 3 * (J - 1); ++i) smpl[i] = s[i];
}

}  // extern "C"
