// oracle/ref_avatar.cpp -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own Avatar.cpp and GaussianMixture.cpp,
// which oracle/Makefile compiles from /root/reference (never copied) against the Eigen stand-in under oracle/shim.  The
// tests use it to check that the oracle's restatement of Avatar::update (Avatar.cpp:22-75), GaussianMixture::load / residual /
// pdf (GaussianMixture.cpp:12-114), Avatar::smplParams (:128-137) and Avatar::alignToJoints (:141-193) computes what the
// reference's code computes.  The model is built by the reference's own loader: AvatarModel.cpp (+ its vendored cnpy.cpp)
// reads <dir>/model.npz and pose_prior.txt; ref_model_tables hands out what it derived so that the oracle's restatement of the
// loader (assigned joints, joint shape regressor, mesh, kinematic tree) can be compared with it.
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "Avatar.h"
#include "GaussianMixture.h"
#include "Util.h"

namespace {
struct RefModel {   // an ark::AvatarModel built by the reference's own constructor (AvatarModel.cpp) from a model directory
    ark::AvatarModel m;
    explicit RefModel(const char* dir) : m(dir) {}
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

// ---- AvatarModel through the reference's own loader: <dir>/model.npz (+ pose_prior.txt) ----
void* ref_model_load(const char* dir) { return new RefModel(dir); }
void ref_model_dims(void* h, int32_t* out8) {
    ark::AvatarModel& m = static_cast<RefModel*>(h)->m;
    int total = 0;
    for (auto& a : m.assignedJoints) total += (int)a.size();
    out8[0] = m.numPoints(); out8[1] = m.numJoints(); out8[2] = m.numShapeKeys(); out8[3] = m.numFaces();
    out8[4] = m.posePrior.nComps; out8[5] = m.posePrior.nDims; out8[6] = m.useJointShapeRegressor ? 1 : 0; out8[7] = total;
}
// the tables the loader derives: parent [J], baseCloud [3V], keyClouds [3V][K] row-major, jointShapeRegBase [3J],
// jointShapeReg [3J][K] row-major, initialJointPos [J][3], mesh [F][3], assigned CSR (start [V+1], joint, weight)
void ref_model_tables(void* h, int32_t* parent, double* base, double* key, double* jsr_base, double* jsr, double* init_pos,
                      int32_t* mesh, int32_t* asg_start, int32_t* asg_joint, double* asg_weight) {
    ark::AvatarModel& m = static_cast<RefModel*>(h)->m;
    const int V = m.numPoints(), J = m.numJoints(), K = m.numShapeKeys(), F = m.numFaces();
    for (int j = 0; j < J; ++j) parent[j] = m.parent[j];
    for (int i = 0; i < 3 * V; ++i) {
        base[i] = m.baseCloud[i];
        for (int k = 0; k < K; ++k) key[(size_t)i * K + k] = m.keyClouds(i, k);
    }
    for (int i = 0; i < 3 * J; ++i) {
        jsr_base[i] = m.jointShapeRegBase[i];
        for (int k = 0; k < K; ++k) jsr[(size_t)i * K + k] = m.jointShapeReg(i, k);
    }
    for (int j = 0; j < J; ++j)
        for (int c = 0; c < 3; ++c) init_pos[3 * j + c] = m.initialJointPos(c, j);
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c) mesh[3 * f + c] = m.mesh(c, f);
    int e = 0;
    for (int v = 0; v < V; ++v) {
        asg_start[v] = e;
        for (auto& wj : m.assignedJoints[(size_t)v]) {
            asg_weight[e] = wj.first;
            asg_joint[e] = wj.second;
            ++e;
        }
    }
    asg_start[V] = e;
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
