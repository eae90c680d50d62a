/* oracle/avatar_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see avatar_oracle.h).
 *
 * CPU restatement of the sxyu/avatar fitting path.  Parity status: forward model / visibility /
 * NN / residuals / Jacobians / priors follow the reference source line by line (citations at
 * each function, relative to /root/reference) and are PINNED against the reference's own code:
 * Avatar.cpp, GaussianMixture.cpp, AvatarOptimizer.cpp and AvatarRenderer.cpp compile unchanged into
 * oracle/_ref/libref_avatar.so (stand-in Eigen / Ceres / boost / OpenCV headers in oracle/shim) and
 * tests/test_oracle.py compares cost, gradient, J^T J (1e-15) and whole fits (1e-12) with them.
 * The solver POLICY (Ceres 1.14, not vendored, not installed) is restated from its published
 * algorithm and stays PARITY UNPINNED.
 */
#include "avatar_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

using V3 = std::array<double, 3>;
struct M3 {
    double m[9];  // row-major
    double& operator()(int r, int c) { return m[3 * r + c]; }
    double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline M3 m3_identity() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 mul(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}
inline V3 mul(const M3& a, const V3& v) {
    return V3{a(0, 0) * v[0] + a(0, 1) * v[1] + a(0, 2) * v[2],
              a(1, 0) * v[0] + a(1, 1) * v[1] + a(1, 2) * v[2],
              a(2, 0) * v[0] + a(2, 1) * v[1] + a(2, 2) * v[2]};
}
inline V3 add(const V3& a, const V3& b) { return V3{a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline V3 sub(const V3& a, const V3& b) { return V3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline V3 scale(const V3& a, double s) { return V3{a[0] * s, a[1] * s, a[2] * s}; }

struct Quat {  // Eigen::Quaterniond coeffs order (x, y, z, w), SURVEY Appendix A
    double x, y, z, w;
};

// Eigen Quaterniond::toRotationMatrix (standard formula, no normalisation) -- SURVEY Appendix A
M3 quat_to_rot(const Quat& q) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
}

// Eigen Quaterniond(Matrix3d) (Shepperd branches) -- SURVEY Appendix A
Quat rot_to_quat_raw(const M3& m) {
    double q[4];  // x y z w
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t;
        q[1] = (m(0, 2) - m(2, 0)) * t;
        q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t;
        q[j] = (m(j, i) + m(i, j)) * t;
        q[k] = (m(k, i) + m(i, k)) * t;
    }
    return Quat{q[0], q[1], q[2], q[3]};
}

// Eigen AngleAxisd(Quaterniond): angle in [0, pi] -- SURVEY Appendix A
void quat_to_angle_axis(const Quat& q, double& angle, V3& axis) {
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (n != 0) {
        angle = 2 * std::atan2(n, std::fabs(q.w));
        if (q.w < 0) n = -n;
        axis = V3{q.x / n, q.y / n, q.z / n};
    } else {
        angle = 0;
        axis = V3{1, 0, 0};
    }
}
Quat angle_axis_to_quat(double angle, const V3& axis) {
    const double ha = 0.5 * angle, s = std::sin(ha);
    return Quat{s * axis[0], s * axis[1], s * axis[2], std::cos(ha)};
}
// Eigen quaternion product a * b
Quat qmul(const Quat& a, const Quat& b) {
    return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
                a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}

// lower Cholesky A = L L^T of an n x n row-major SPD matrix (Eigen LLT semantics); false if not PD
bool cholesky_lower(const double* A, int n, double* L) {
    std::fill(L, L + (size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        if (!(d > 0)) return false;
        d = std::sqrt(d);
        L[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
            L[(size_t)i * n + j] = s / d;
        }
    }
    return true;
}
// solve L L^T x = b in place
void cholesky_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
        b[i] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
        b[i] = s / L[(size_t)i * n + i];
    }
}

/* ---------------- GaussianMixture (GaussianMixture.cpp) ---------------- */
struct GMM {
    int nComps = -1, nDims = 0;
    std::vector<double> weight, mean, cov, cov_cho, prec_cho, consts, consts_log;

    // GaussianMixture.cpp:22-76 on parsed arrays
    int finish_load() {
        const int C = nComps, D = nDims;
        const double sqrt_2_pi_n = std::pow(2 * M_PI, D * 0.5);
        const double log_sqrt_2_pi_n = D * 0.5 * std::log(2 * M_PI);
        consts.assign(C, 0);
        consts_log.assign(C, 0);
        for (int i = 0; i < C; ++i) {
            consts_log[i] = std::log(weight[i]) - log_sqrt_2_pi_n;
            consts[i] = weight[i] / sqrt_2_pi_n;
        }
        cov_cho.assign((size_t)C * D * D, 0);
        prec_cho.assign((size_t)C * D * D, 0);
        double minDet = std::numeric_limits<double>::max();
        std::vector<double> inv((size_t)D * D), col(D);
        for (int i = 0; i < C; ++i) {
            const double* S = &cov[(size_t)i * D * D];
            double* L = &cov_cho[(size_t)i * D * D];
            if (!cholesky_lower(S, D, L)) return 1;  // "Decomposition failed!" (:60)
            // cov.inverse() (:61) -- here via the Cholesky factor, column by column
            for (int c = 0; c < D; ++c) {
                std::fill(col.begin(), col.end(), 0.0);
                col[c] = 1.0;
                cholesky_solve(L, D, col.data());
                for (int r = 0; r < D; ++r) inv[(size_t)r * D + c] = col[r];
            }
            for (int r = 0; r < D; ++r)
                for (int c = r + 1; c < D; ++c) {
                    const double s = 0.5 * (inv[(size_t)r * D + c] + inv[(size_t)c * D + r]);
                    inv[(size_t)r * D + c] = inv[(size_t)c * D + r] = s;
                }
            if (!cholesky_lower(inv.data(), D, &prec_cho[(size_t)i * D * D])) return 1;
            double det = 1.0;  // determinant of the triangular factor (:63)
            for (int d = 0; d < D; ++d) det *= L[(size_t)d * D + d];
            minDet = std::min(det, minDet);
            consts[i] /= det;
            consts_log[i] -= std::log(det);
        }
        for (int i = 0; i < C; ++i) {
            consts[i] *= minDet;
            consts_log[i] += std::log(minDet);
        }
        return 0;
    }

    // GaussianMixture.cpp:12-77 (text format)
    int load_text(const std::string& path) {
        FILE* fp = std::fopen(path.c_str(), "r");
        if (!fp) {
            nComps = -1;
            return 0;
        }
        bool ok = std::fscanf(fp, "%d %d", &nComps, &nDims) == 2;
        const int C = nComps, D = nDims;
        auto read_all = [&](std::vector<double>& v, size_t n) {
            v.resize(n);
            for (size_t i = 0; i < n && ok; ++i) ok = std::fscanf(fp, "%lf", &v[i]) == 1;
        };
        if (ok) read_all(weight, C);
        if (ok) read_all(mean, (size_t)C * D);
        if (ok) read_all(cov, (size_t)C * D * D);
        std::fclose(fp);
        if (!ok) return 2;
        return finish_load();
    }

    // GaussianMixture.cpp:95-114; out has nDims+1 entries
    int residual(const double* x, double* out) const {
        const int D = nDims;
        double bestProb = std::numeric_limits<double>::max();
        int best = -1;
        std::vector<double> res(D + 1), diff(D);
        for (int i = 0; i < nComps; ++i) {
            const double* L = &prec_cho[(size_t)i * D * D];
            for (int d = 0; d < D; ++d) diff[d] = x[d] - mean[(size_t)i * D + d];
            res[D] = 0.;
            double sq = 0;
            for (int c = 0; c < D; ++c) {  // (L^T diff)_c = sum_{r>=c} L(r,c) diff_r
                double s = 0;
                for (int r = c; r < D; ++r) s += L[(size_t)r * D + c] * diff[r];
                res[c] = s * std::sqrt(0.5);
                sq += res[c] * res[c];
            }
            const double p = sq - consts_log[i];
            if (p < bestProb) {
                bestProb = p;
                res[D] = std::sqrt(-consts_log[i]);
                std::copy(res.begin(), res.end(), out);
                best = i;
            }
        }
        return best;
    }
};

}  // namespace

/* ---------------- AvatarModel (AvatarModel.cpp:23-127, include/Avatar.h:64-151) ---------------- */
struct orc_model {
    int V = 0, J = 0, K = 0, F = 0;
    std::vector<double> baseCloud;          // 3V
    std::vector<double> keyClouds;          // (3V) x K, [3v+c][k]
    std::vector<double> initialJointPos;    // 3J
    std::vector<double> jointShapeRegBase;  // 3J
    std::vector<double> jointShapeReg;      // (3J) x K
    std::vector<int> parent;
    std::vector<int> mesh;                  // 3F
    std::vector<std::vector<std::pair<double, int>>> assignedJoints;  // (weight, joint) desc
    std::vector<std::vector<std::pair<int, double>>> weightsCol;      // sparse weights column v: (joint asc, w)
    GMM posePrior;
};

struct orc_optimizer {
    const orc_model* model;
    int numParts;
    std::vector<int> partMap;
    std::vector<std::vector<int>> modelPartIndices;  // AvatarOptimizer.cpp:1223-1243
    std::vector<int> modelPartLabels;                // :1307-1311
};

namespace {

/* ---------------- Avatar::update (Avatar.cpp:22-75) ---------------- */
void avatar_update(const orc_model& md, const double* p, const M3* R, const double* w,
                   double* cloud, double* jointPosOut, double* jointTransOut) {
    const int V = md.V, J = md.J, K = md.K;
    std::vector<double> shaped(3 * (size_t)V);
    for (int r = 0; r < 3 * V; ++r) {  // :26
        double s = 0;
        for (int k = 0; k < K; ++k) s += md.keyClouds[(size_t)r * K + k] * w[k];
        shaped[r] = s + md.baseCloud[r];
    }
    std::vector<double> jointPos(md.initialJointPos);  // :33 (useJointShapeRegressor branch)
    for (int r = 0; r < 3 * J; ++r) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += md.jointShapeReg[(size_t)r * K + k] * w[k];
        jointPos[r] += s;
    }
    struct Aff { M3 L; V3 t; };
    std::vector<Aff> jt(J);
    jt[0].L = R[0];  // :47-49
    jt[0].t = V3{p[0], p[1], p[2]};
    for (int i = 1; i < J; ++i) {  // :50-57, util::mulAffine (Util.h:191-199)
        const int pa = md.parent[i];
        V3 rel{jointPos[3 * i] - jointPos[3 * pa], jointPos[3 * i + 1] - jointPos[3 * pa + 1],
               jointPos[3 * i + 2] - jointPos[3 * pa + 2]};
        jt[i].L = mul(jt[pa].L, R[i]);
        jt[i].t = add(jt[pa].t, mul(jt[pa].L, rel));
    }
    for (int i = 0; i < J; ++i) {  // :59-64
        V3 jInit{jointPos[3 * i], jointPos[3 * i + 1], jointPos[3 * i + 2]};
        for (int c = 0; c < 3; ++c) jointPos[3 * i + c] = jt[i].t[c];
        jt[i].t = sub(jt[i].t, mul(jt[i].L, jInit));
    }
    for (int v = 0; v < V; ++v) {  // :69-73: pointTrans = jointTrans * weights, then apply
        double T[12] = {0};         // 3x4 column-major
        for (auto& jw : md.weightsCol[v]) {
            const Aff& a = jt[jw.first];
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) T[3 * c + r] += a.L(r, c) * jw.second;
            for (int r = 0; r < 3; ++r) T[9 + r] += a.t[r] * jw.second;
        }
        const double* s = &shaped[3 * (size_t)v];
        for (int r = 0; r < 3; ++r)
            cloud[3 * (size_t)v + r] = T[r] * s[0] + T[3 + r] * s[1] + T[6 + r] * s[2] + T[9 + r];
    }
    if (jointPosOut) std::copy(jointPos.begin(), jointPos.end(), jointPosOut);
    if (jointTransOut)
        for (int i = 0; i < J; ++i) {
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) jointTransOut[12 * i + 3 * c + r] = jt[i].L(r, c);
            for (int r = 0; r < 3; ++r) jointTransOut[12 * i + 9 + r] = jt[i].t[r];
        }
}

/* ---------------- AvatarEvaluationCommonData + AvatarCostFunctorCache ---------------- */
struct Ancestor {  // AvatarOptimizer.cpp:368-404
    int jid;
    int assign[4];
    double weight[4];
    int num_assign;
};

struct Cache {  // AvatarOptimizer.cpp:459-605
    int pointId;
    V3 resid;
    std::vector<M3> icpJacobian;          // per ancestor, 3x3 (already times localJacobian)
    std::vector<double> icpShapeJacobian;  // 3 x K row-major
};

struct Common {
    const orc_model& md;
    const int J, K, V, nJS;
    std::vector<std::vector<Ancestor>> ancestor;
    std::vector<double> S, Sp, H;  // J x (3 x K row-major)
    std::vector<double> shapedCloud, jointPosInit, jointVecInit;
    std::vector<M3> R_;
    std::vector<V3> t_;
    std::vector<std::array<double, 12>> localJacobian;  // 4x3 row-major
    std::vector<Cache> caches;
    // current parameters (the buffers Ceres optimises in place, :1407-1413)
    V3 p;
    std::vector<Quat> q;
    std::vector<double> w;
    int numThreads = 1;
    double scaledBetaPose = 0, scaledBetaShape = 0;

    M3& R(int a, int j) { return R_[(size_t)nJS * (a + 1) + j + 1]; }  // :350-352
    V3& t(int a, int j) { return t_[(size_t)nJS * (a + 1) + j + 1]; }  // :354-356

    explicit Common(const orc_model& m)  // :169-246
        : md(m), J(m.J), K(m.K), V(m.V), nJS(m.J + 1) {
        localJacobian.resize(nJS);
        S.assign((size_t)J * 3 * K, 0);
        Sp.assign((size_t)J * 3 * K, 0);
        H.assign((size_t)J * 3 * K, 0);
        R_.resize((size_t)nJS * nJS);
        t_.resize(R_.size());
        shapedCloud.resize(3 * (size_t)V);
        jointPosInit.resize(3 * (size_t)J);
        jointVecInit.resize(3 * (size_t)J);
        q.resize(J);
        w.assign(K, 0);
        ancestor.resize(V);
        for (int point = 0; point < V; ++point) {  // :189-213
            std::vector<Ancestor> ances;
            for (auto& wj : md.assignedJoints[point]) {
                const double weight = wj.first;
                const int joint = wj.second;
                auto mk = [&](int jid) {
                    Ancestor a;
                    a.jid = jid;
                    a.assign[0] = joint;
                    a.weight[0] = weight;
                    a.num_assign = 1;
                    return a;
                };
                ances.push_back(mk(joint));
                for (int j = md.parent[joint]; j != -1; j = md.parent[j]) ances.push_back(mk(j));
            }
            // :200 uses std::sort (unstable); stable here so that the merge order is defined
            std::stable_sort(ances.begin(), ances.end(),
                             [](const Ancestor& a, const Ancestor& b) { return a.jid < b.jid; });
            size_t last = 0;
            for (size_t i = 1; i < ances.size(); ++i) {
                if (ances[last].jid == ances[i].jid) {
                    for (int k = 0; k < ances[i].num_assign; ++k) {  // merge (:398-403)
                        ances[last].weight[ances[last].num_assign] = ances[i].weight[k];
                        ances[last].assign[ances[last].num_assign++] = ances[i].assign[k];
                    }
                } else {
                    ++last;
                    if (last < i) ances[last] = ances[i];
                }
            }
            ances.resize(last + 1);
            ancestor[point] = std::move(ances);
        }
        for (int j = 0; j < J; ++j)  // :222-226
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < K; ++k)
                    S[((size_t)j * 3 + r) * K + k] = md.jointShapeReg[((size_t)3 * j + r) * K + k];
        for (int j = 1; j < J; ++j)  // :240-242
            for (int e = 0; e < 3 * K; ++e)
                Sp[(size_t)j * 3 * K + e] = S[(size_t)j * 3 * K + e] - S[(size_t)md.parent[j] * 3 * K + e];
    }

    void CalcShape() {  // :249-281
        for (int r = 0; r < 3 * V; ++r) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += md.keyClouds[(size_t)r * K + k] * w[k];
            shapedCloud[r] = s + md.baseCloud[r];
        }
        for (int r = 0; r < 3 * J; ++r) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += md.jointShapeReg[(size_t)r * K + k] * w[k];
            jointPosInit[r] = md.jointShapeRegBase[r] + s;
        }
        const V3 offset{jointPosInit[0], jointPosInit[1], jointPosInit[2]};
        for (int v = 0; v < V; ++v)
            for (int c = 0; c < 3; ++c) shapedCloud[3 * (size_t)v + c] -= offset[c];
        for (int j = 0; j < J; ++j)
            for (int c = 0; c < 3; ++c) jointPosInit[3 * j + c] -= offset[c];
        jointVecInit = jointPosInit;
        for (int i = J - 1; i >= 1; --i)
            for (int c = 0; c < 3; ++c) jointVecInit[3 * i + c] -= jointVecInit[3 * md.parent[i] + c];
    }

    void updateData(Cache& ch, bool compute_jacobians) {  // :505-582
        const int pointId = ch.pointId;
        const V3 pointPosInit{shapedCloud[3 * (size_t)pointId], shapedCloud[3 * (size_t)pointId + 1],
                              shapedCloud[3 * (size_t)pointId + 2]};
        auto jp = [&](int k) { return V3{jointPosInit[3 * k], jointPosInit[3 * k + 1], jointPosInit[3 * k + 2]}; };
        ch.resid = V3{0, 0, 0};
        for (auto& assign : md.assignedJoints[pointId]) {
            const int k = assign.second;
            ch.resid = add(ch.resid, scale(add(mul(R(-1, k), sub(pointPosInit, jp(k))), t(-1, k)), assign.first));
        }
        if (!compute_jacobians) return;
        const auto& anc = ancestor[pointId];
        for (size_t i = 0; i < anc.size(); ++i) {
            const Ancestor& ances = anc[i];
            const int j = ances.jid;
            V3 v{0, 0, 0};
            for (int a = 0; a < ances.num_assign; ++a) {
                const int k = ances.assign[a];
                v = add(v, scale(add(mul(R(j, k), sub(pointPosInit, jp(k))), t(j, k)), ances.weight[a]));
            }
            const Quat& qq = q[j];
            const double u[3] = {qq.x * 2, qq.y * 2, qq.z * 2};
            const double ww = qq.w * 2;
            double dRot[12];  // 3x4 row-major (:546-558)
            dRot[0] = u[1] * v[1] + v[2] * u[2];
            dRot[1] = ww * v[2] + u[0] * v[1] - 2 * u[1] * v[0];
            dRot[2] = -ww * v[1] - 2 * v[0] * u[2] + u[0] * v[2];
            dRot[3] = u[1] * v[2] - v[1] * u[2];
            dRot[4] = -ww * v[2] - 2 * u[0] * v[1] + v[0] * u[1];
            dRot[5] = v[2] * u[2] + u[0] * v[0];
            dRot[6] = ww * v[0] + u[1] * v[2] - 2 * v[1] * u[2];
            dRot[7] = v[0] * u[2] - u[0] * v[2];
            dRot[8] = ww * v[1] + v[0] * u[2] - 2 * u[0] * v[2];
            dRot[9] = -ww * v[0] - 2 * u[1] * v[2] + v[1] * u[2];
            dRot[10] = u[0] * v[0] + v[1] * u[1];
            dRot[11] = u[0] * v[1] - v[0] * u[1];
            const M3& Rp = R(-1, md.parent[j]);
            const auto& Lq = localJacobian[j];
            double RD[12];  // Rp * dRot, 3x4
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c)
                    RD[4 * r + c] = Rp(r, 0) * dRot[c] + Rp(r, 1) * dRot[4 + c] + Rp(r, 2) * dRot[8 + c];
            M3 out;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    out(r, c) = RD[4 * r] * Lq[c] + RD[4 * r + 1] * Lq[3 + c] + RD[4 * r + 2] * Lq[6 + c] +
                                RD[4 * r + 3] * Lq[9 + c];
            ch.icpJacobian[i] = out;
        }
        std::fill(ch.icpShapeJacobian.begin(), ch.icpShapeJacobian.end(), 0.0);  // :568-580
        for (auto& assign : md.assignedJoints[pointId]) {
            const int j = assign.second;
            const M3& Rj = R(-1, j);
            const double* Sj = &S[(size_t)j * 3 * K];
            const double* Hj = &H[(size_t)j * 3 * K];
            const double* pd = &md.keyClouds[(size_t)pointId * 3 * K];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < K; ++k) {
                    const double val = Rj(r, 0) * (pd[k] - Sj[k]) + Rj(r, 1) * (pd[K + k] - Sj[K + k]) +
                                       Rj(r, 2) * (pd[2 * K + k] - Sj[2 * K + k]) + Hj[r * K + k];
                    ch.icpShapeJacobian[r * K + k] += val * assign.first;
                }
        }
    }

    void PrepareForEvaluation(bool evaluate_jacobians) {  // :283-347 (new_evaluation_point = true)
        CalcShape();
        for (int c = 0; c < 3; ++c) jointVecInit[c] = p[c];  // :291
        for (int i = 0; i < J; ++i) {                          // :293-299
            const Quat& x = q[i];
            localJacobian[i] = {x.w, x.z, -x.y, -x.z, x.w, x.x, x.y, -x.x, x.w, -x.x, -x.y, -x.z};
        }
        R(-1, -1) = m3_identity();  // :303-315
        t(-1, -1) = V3{0, 0, 0};
        for (int i = 0; i < J; ++i) {
            R(i, i) = m3_identity();
            const M3 rot = quat_to_rot(q[i]);
            t(i, i) = V3{0, 0, 0};
            const int pa = md.parent[i];
            const V3 jv{jointVecInit[3 * i], jointVecInit[3 * i + 1], jointVecInit[3 * i + 2]};
            for (int j = pa;; j = md.parent[j]) {
                R(j, i) = mul(R(j, pa), rot);
                t(j, i) = add(mul(R(j, pa), jv), t(j, pa));
                if (j == -1) break;
            }
        }
        for (int j = 1; j < J; ++j) {  // :318-324
            const M3& Rp = R(-1, md.parent[j]);
            const double* Spj = &Sp[(size_t)j * 3 * K];
            const double* Hp = &H[(size_t)md.parent[j] * 3 * K];
            double* Hj = &H[(size_t)j * 3 * K];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < K; ++k)
                    Hj[r * K + k] = Rp(r, 0) * Spj[k] + Rp(r, 1) * Spj[K + k] + Rp(r, 2) * Spj[2 * K + k] + Hp[r * K + k];
        }
        if (numThreads <= 1 || caches.size() < 64) {
            for (auto& c : caches) updateData(c, evaluate_jacobians);
        } else {  // :327-343: spawn + join num_threads std::threads per evaluation, atomic counter
            std::atomic<size_t> cacheId(0);
            auto worker = [&]() {
                while (true) {
                    size_t id = cacheId++;
                    if (id >= caches.size()) break;
                    updateData(caches[id], evaluate_jacobians);
                }
            };
            std::vector<std::thread> pool;
            for (int i = 0; i < numThreads; ++i) pool.emplace_back(worker);
            for (auto& th : pool) th.join();
        }
    }

    void set_params(const double* x) {
        p = V3{x[0], x[1], x[2]};
        for (int j = 0; j < J; ++j) q[j] = Quat{x[3 + 4 * j], x[4 + 4 * j], x[5 + 4 * j], x[6 + 4 * j]};
        for (int k = 0; k < K; ++k) w[k] = x[3 + 4 * J + k];
    }
    void get_params(double* x) const {
        for (int c = 0; c < 3; ++c) x[c] = p[c];
        for (int j = 0; j < J; ++j) {
            x[3 + 4 * j] = q[j].x; x[4 + 4 * j] = q[j].y; x[5 + 4 * j] = q[j].z; x[6 + 4 * j] = q[j].w;
        }
        for (int k = 0; k < K; ++k) x[3 + 4 * J + k] = w[k];
    }
    void make_cache(int pointId) {  // :460-471
        Cache c;
        c.pointId = pointId;
        c.icpJacobian.resize(ancestor[pointId].size());
        c.icpShapeJacobian.assign(3 * (size_t)K, 0);
        caches.push_back(std::move(c));
    }
};

/* The Ceres problem of one ICP iteration: residual blocks in AddResidualBlock order
 * (AvatarOptimizer.cpp:1442-1474) and the cost / tangent gradient / GN matrix Ceres derives. */
struct Problem {
    Common& cm;
    const double* data;
    std::vector<std::vector<int>> corr;  // per cache: data point ids
    size_t totalResiduals = 0;
    double betaPose, betaShape;
    int P;
    long evaluations = 0;

    Problem(Common& c, const double* d, double bp, double bs)
        : cm(c), data(d), betaPose(bp), betaShape(bs), P(3 + 3 * c.J + c.K) {}

    // cost = 1/2 sum |r|^2; grad = J_local^T r (FakeQuaternionParameterization::ComputeJacobian = [I;0], :144-149)
    double evaluate(const double* x, double* grad, double* Hgn) {
        ++evaluations;
        const int J = cm.J, K = cm.K;
        cm.set_params(x);
        const bool jac = (grad != nullptr) || (Hgn != nullptr);
        cm.PrepareForEvaluation(jac);
        if (grad) std::fill(grad, grad + P, 0.0);
        if (Hgn) std::fill(Hgn, Hgn + (size_t)P * P, 0.0);
        double cost = 0;
        std::vector<int> cols;
        std::vector<double> Jrow;  // 3 x ncols
        for (size_t ci = 0; ci < cm.caches.size(); ++ci) {
            const Cache& ch = cm.caches[ci];
            const auto& anc = cm.ancestor[ch.pointId];
            if (Hgn) {  // dense row block for this vertex: p | ancestors | w   (:473-503)
                cols.clear();
                for (int c = 0; c < 3; ++c) cols.push_back(c);
                for (auto& a : anc)
                    for (int c = 0; c < 3; ++c) cols.push_back(3 + 3 * a.jid + c);
                for (int k = 0; k < K; ++k) cols.push_back(3 + 3 * J + k);
                const int L = (int)cols.size();
                Jrow.assign(3 * (size_t)L, 0.0);
                for (int r = 0; r < 3; ++r) {
                    Jrow[(size_t)r * L + r] = 1.0;
                    for (size_t a = 0; a < anc.size(); ++a)
                        for (int c = 0; c < 3; ++c) Jrow[(size_t)r * L + 3 + 3 * a + c] = ch.icpJacobian[a](r, c);
                    for (int k = 0; k < K; ++k) Jrow[(size_t)r * L + 3 + 3 * anc.size() + k] = ch.icpShapeJacobian[r * K + k];
                }
            }
            for (int di : corr[ci]) {  // AvatarICPCostFunctor::Evaluate (:632-639)
                const double r[3] = {ch.resid[0] - data[3 * (size_t)di], ch.resid[1] - data[3 * (size_t)di + 1],
                                     ch.resid[2] - data[3 * (size_t)di + 2]};
                cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                if (grad) {
                    for (int c = 0; c < 3; ++c) grad[c] += r[c];
                    for (size_t a = 0; a < anc.size(); ++a) {
                        const M3& B = ch.icpJacobian[a];
                        double* g = grad + 3 + 3 * anc[a].jid;
                        for (int c = 0; c < 3; ++c) g[c] += B(0, c) * r[0] + B(1, c) * r[1] + B(2, c) * r[2];
                    }
                    double* gw = grad + 3 + 3 * J;
                    const double* SJ = ch.icpShapeJacobian.data();
                    for (int k = 0; k < K; ++k) gw[k] += SJ[k] * r[0] + SJ[K + k] * r[1] + SJ[2 * K + k] * r[2];
                }
                if (Hgn) {
                    const int L = (int)cols.size();
                    for (int a = 0; a < L; ++a) {
                        const double j0 = Jrow[a], j1 = Jrow[L + a], j2 = Jrow[2 * (size_t)L + a];
                        double* Hr = Hgn + (size_t)cols[a] * P;
                        for (int b = 0; b < L; ++b)
                            Hr[cols[b]] += j0 * Jrow[b] + j1 * Jrow[L + b] + j2 * Jrow[2 * (size_t)L + b];
                    }
                }
            }
        }
        // AvatarPosePriorCostFunctor::Evaluate (:661-692)
        if (betaPose > 0.) {
            const GMM& g = cm.md.posePrior;
            const int n = J - 1, D = 3 * n;
            std::vector<double> smpl(D), res(D + 1);
            for (int i = 0; i < n; ++i) {
                double ang; V3 ax;
                quat_to_angle_axis(cm.q[i + 1], ang, ax);
                for (int c = 0; c < 3; ++c) smpl[3 * i + c] = ax[c] * ang;
            }
            const int comp = g.residual(smpl.data(), res.data());
            double sq = 0;
            for (int d = 0; d <= D; ++d) {
                res[d] *= cm.scaledBetaPose;
                sq += res[d] * res[d];
            }
            cost += 0.5 * sq;
            if (jac) {
                const double* L = &g.prec_cho[(size_t)comp * D * D];
                const double f = 0.707106781186548 * cm.scaledBetaPose;  // :684-685
                // J(:, 3i..3i+2) = L.middleRows<3>(3i)^T * f  => J = f * L^T (D x D); last row zero
                if (grad)
                    for (int r = 0; r < D; ++r) {  // grad_r = f * sum_c L(r,c) res_c, c <= r
                        double s = 0;
                        for (int c = 0; c <= r; ++c) s += L[(size_t)r * D + c] * res[c];
                        grad[6 + r] += f * s;
                    }
                if (Hgn)
                    for (int a = 0; a < D; ++a)
                        for (int b = 0; b < D; ++b) {
                            double s = 0;
                            const int m = std::min(a, b);
                            for (int c = 0; c <= m; ++c) s += L[(size_t)a * D + c] * L[(size_t)b * D + c];
                            Hgn[(size_t)(6 + a) * P + 6 + b] += f * f * s;
                        }
            }
        }
        // AvatarShapePriorCostFunctor::Evaluate (:708-723)
        if (betaShape > 0.) {
            for (int k = 0; k < K; ++k) {
                const double r = cm.w[k] * cm.scaledBetaShape;
                cost += 0.5 * r * r;
                if (grad) grad[3 + 3 * J + k] += cm.scaledBetaShape * r;
                if (Hgn) Hgn[(size_t)(3 + 3 * J + k) * P + 3 + 3 * J + k] += cm.scaledBetaShape * cm.scaledBetaShape;
            }
        }
        return cost;
    }

    // FakeQuaternionParameterization::Plus for q (:123-143); p, w additive
    void plus(const double* x, const double* delta, double* xp) const {
        const int J = cm.J, K = cm.K;
        for (int c = 0; c < 3; ++c) xp[c] = x[c] + delta[c];
        for (int j = 0; j < J; ++j) {
            const double* d = delta + 3 + 3 * j;
            const Quat qx{x[3 + 4 * j], x[4 + 4 * j], x[5 + 4 * j], x[6 + 4 * j]};
            const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            Quat r = qx;
            if (nd > 0.0) {
                const double s = std::sin(nd) / nd;
                r = qmul(Quat{s * d[0], s * d[1], s * d[2], std::cos(nd)}, qx);
            }
            xp[3 + 4 * j] = r.x; xp[4 + 4 * j] = r.y; xp[5 + 4 * j] = r.z; xp[6 + 4 * j] = r.w;
        }
        for (int k = 0; k < K; ++k) xp[3 + 4 * J + k] = x[3 + 4 * J + k] + delta[3 + 3 * J + k];
    }
};

/* ---------------- visibility (AvatarOptimizer.cpp:1349-1367) ---------------- */
void visibility(const orc_model& md, const double* cloud, uint8_t* vis) {
    std::fill(vis, vis + md.V, (uint8_t)0);
    for (int f = 0; f < md.F; ++f) {
        const int i1 = md.mesh[3 * f], i2 = md.mesh[3 * f + 1], i3 = md.mesh[3 * f + 2];
        const double* p1 = cloud + 3 * (size_t)i1;
        const double* p2 = cloud + 3 * (size_t)i2;
        const double* p3 = cloud + 3 * (size_t)i3;
        // ((p2 - p1).cross(p1 - p3)).z()
        const double ax = p2[0] - p1[0], ay = p2[1] - p1[1];
        const double bx = p1[0] - p3[0], by = p1[1] - p3[1];
        if (ax * by - ay * bx > 1e-4) vis[i1] = vis[i2] = vis[i3] = 1;
    }
}

/* ---------------- exact 1-NN (nanoflann semantics: L2_Simple, eps 0, strict <) ----------------
 * nanoflann.hpp:423-445 (distance = sum_c (a_c - b_c)^2, c = 0,1,2 in order), :142-205 (result set:
 * accept iff dist < worst).  Tie rule here: lowest compacted index wins (nanoflann: first visited). */
inline double dist2(const double* a, const double* b) {
    double s = 0;
    for (int c = 0; c < 3; ++c) {
        const double d = a[c] - b[c];
        s += d * d;
    }
    return s;
}
struct KdTree {  // own exact kd-tree (leaf 10), used as the oracle's fast path
    struct Node { int lo, hi, dim; double split; int left, right; };
    const double* pts;
    std::vector<int> idx;
    std::vector<Node> nodes;
    KdTree(const double* p, int n) : pts(p), idx(n) {
        std::iota(idx.begin(), idx.end(), 0);
        nodes.reserve(2 * n / 10 + 4);
        build(0, n);
    }
    int build(int lo, int hi) {
        const int id = (int)nodes.size();
        nodes.push_back(Node{lo, hi, -1, 0, -1, -1});
        if (hi - lo <= 10) return id;
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = lo; i < hi; ++i)
            for (int c = 0; c < 3; ++c) {
                mn[c] = std::min(mn[c], pts[3 * (size_t)idx[i] + c]);
                mx[c] = std::max(mx[c], pts[3 * (size_t)idx[i] + c]);
            }
        int dim = 0;
        for (int c = 1; c < 3; ++c)
            if (mx[c] - mn[c] > mx[dim] - mn[dim]) dim = c;
        const int mid = (lo + hi) / 2;
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
            return pts[3 * (size_t)a + dim] < pts[3 * (size_t)b + dim];
        });
        const double split = pts[3 * (size_t)idx[mid] + dim];
        const int l = build(lo, mid);
        const int r = build(mid, hi);
        nodes[id].dim = dim;
        nodes[id].split = split;
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }
    void search(int id, const double* qp, double& best, int& bi) const {
        const Node& n = nodes[id];
        if (n.dim < 0) {
            for (int i = n.lo; i < n.hi; ++i) {
                const double d = dist2(qp, pts + 3 * (size_t)idx[i]);
                if (d < best || (d == best && idx[i] < bi)) {
                    best = d;
                    bi = idx[i];
                }
            }
            return;
        }
        const double diff = qp[n.dim] - n.split;
        const int near = diff < 0 ? n.left : n.right, far = diff < 0 ? n.right : n.left;
        search(near, qp, best, bi);
        if (diff * diff <= best) search(far, qp, best, bi);
    }
    int closest(const double* qp) const {
        double best = std::numeric_limits<double>::max();
        int bi = -1;
        search(0, qp, best, bi);
        return bi;
    }
};

// findNN(..., invert = true) (AvatarOptimizer.cpp:841-920)
void find_nn(const orc_optimizer& opt, const double* cloud, const uint8_t* vis, const double* data,
             const int32_t* labels, int N, int method, int num_threads, int32_t* out) {
    const int numParts = opt.numParts;
    std::vector<std::vector<double>> partCloud(numParts);
    std::vector<std::vector<int>> newIdx(numParts);
    std::vector<std::unique_ptr<KdTree>> kd(numParts);
    auto build_part = [&](int i) {  // :860-880
        for (int k : opt.modelPartIndices[i]) {
            if (!vis[k]) continue;
            partCloud[i].insert(partCloud[i].end(), cloud + 3 * (size_t)k, cloud + 3 * (size_t)k + 3);
            newIdx[i].push_back(k);
        }
        if (method == 1 && !newIdx[i].empty()) kd[i].reset(new KdTree(partCloud[i].data(), (int)newIdx[i].size()));
    };
    if (num_threads > 1) {  // :853-889 (atomic part counter)
        std::atomic<int> part(0);
        auto worker = [&]() {
            while (true) {
                int i = part++;
                if (i >= numParts) break;
                build_part(i);
            }
        };
        std::vector<std::thread> thds;
        for (int i = 0; i < num_threads; ++i) thds.emplace_back(worker);
        for (auto& t : thds) t.join();
    } else {
        for (int i = 0; i < numParts; ++i) build_part(i);
    }
    for (int i = 0; i < N; ++i) {  // :896-904 (serial)
        const int partId = labels[i];
        if (newIdx[partId].empty()) {
            out[i] = -1;
            continue;
        }
        const double* qp = data + 3 * (size_t)i;
        int bi = -1;
        if (method == 1) {
            bi = kd[partId]->closest(qp);
        } else {
            double best = std::numeric_limits<double>::max();
            const int n = (int)newIdx[partId].size();
            const double* pc = partCloud[partId].data();
            for (int k = 0; k < n; ++k) {
                const double d = dist2(qp, pc + 3 * (size_t)k);
                if (d < best) {
                    best = d;
                    bi = k;
                }
            }
        }
        out[i] = newIdx[partId][bi];
    }
}

/* ---------------- solvers ---------------- */
using EvalFn = std::function<double(const double* x, double* grad, double* H)>;
using PlusFn = std::function<void(const double* x, const double* d, double* xp)>;
using TraceFn = std::function<void(const double* x)>;

struct SolveSummary {
    int iterations = 0, accepted = 0;
    double initial_cost = 0, final_cost = 0;
};

/* Gauss-Newton / Levenberg-Marquardt: restatement of what Ceres 1.14's TRUST_REGION minimizer does
 * with LEVENBERG_MARQUARDT + DENSE_NORMAL_CHOLESKY (the configuration the reference sets at
 * AvatarOptimizer.cpp:1314 and comments out at :1315-1319), Ceres defaults otherwise:
 * initial radius 1e4, max radius 1e16, min radius 1e-32, min_relative_decrease 1e-3,
 * min/max LM diagonal 1e-6/1e32, Jacobi scaling 1/(1+sqrt(diag J^T J)).
 * The reference never executes this; it is the algorithm BASELINE.json mandates for the
 * GPU engine, so this is the GPU path's iterate-for-iterate checker.  PARITY UNPINNED vs Ceres. */
SolveSummary solve_gn_lm(int nx, int P, double* x, const EvalFn& eval, const PlusFn& plus, int max_iters,
                         double function_tolerance, const TraceFn& trace) {
    SolveSummary sum;
    std::vector<double> g(P), H((size_t)P * P), gt(P), Ht((size_t)P * P), A((size_t)P * P), L((size_t)P * P),
        delta(P), xt(nx), Hd(P);
    double cost = eval(x, g.data(), H.data());
    sum.initial_cost = cost;
    double radius = 1e4, decrease_factor = 2.0;
    const double kMinDiag = 1e-6, kMaxDiag = 1e32, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    auto gmax = [&](const std::vector<double>& v) {
        double m = 0;
        for (double e : v) m = std::max(m, std::fabs(e));
        return m;
    };
    bool done = gmax(g) <= gradient_tolerance;
    for (int it = 0; it < max_iters && !done; ++it) {
        ++sum.iterations;
        // (H + D) delta = -g, D_jj = clamp(s_j^2 h_jj, 1e-6, 1e32) / (s_j^2 radius), s_j = 1/(1+sqrt(h_jj))
        A = H;
        for (int j = 0; j < P; ++j) {
            const double hjj = H[(size_t)j * P + j];
            const double s = 1.0 / (1.0 + std::sqrt(hjj));
            const double d = std::min(std::max(s * s * hjj, kMinDiag), kMaxDiag);
            A[(size_t)j * P + j] += d / (s * s * radius);
        }
        bool ok = cholesky_lower(A.data(), P, L.data());
        double model_change = 0;
        if (ok) {
            for (int j = 0; j < P; ++j) delta[j] = -g[j];
            cholesky_solve(L.data(), P, delta.data());
            // model_cost_change = -delta^T (g + 1/2 H delta)
            for (int a = 0; a < P; ++a) {
                double s = 0;
                for (int b = 0; b < P; ++b) s += H[(size_t)a * P + b] * delta[b];
                Hd[a] = s;
            }
            for (int a = 0; a < P; ++a) model_change -= delta[a] * (g[a] + 0.5 * Hd[a]);
            ok = model_change > 0 && std::isfinite(model_change);
        }
        bool accepted = false;
        if (ok) {
            plus(x, delta.data(), xt.data());
            const double cost_t = eval(xt.data(), gt.data(), Ht.data());
            const double rho = (cost - cost_t) / model_change;
            if (std::isfinite(cost_t) && rho > 1e-3) {
                accepted = true;
                ++sum.accepted;
                const double cost_change = cost - cost_t;
                double dn = 0, xn = 0;
                for (int i = 0; i < nx; ++i) {
                    dn += (xt[i] - x[i]) * (xt[i] - x[i]);
                    xn += x[i] * x[i];
                }
                std::copy(xt.begin(), xt.end(), x);
                const double cost_old = cost;
                cost = cost_t;
                g.swap(gt);
                H.swap(Ht);
                radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
                decrease_factor = 2.0;
                if (std::sqrt(dn) <= parameter_tolerance * (std::sqrt(xn) + parameter_tolerance)) done = true;
                if (std::fabs(cost_change) <= function_tolerance * cost_old) done = true;
                if (gmax(g) <= gradient_tolerance) done = true;
            }
        }
        if (!accepted) {
            radius /= decrease_factor;
            decrease_factor *= 2.0;
            if (radius < 1e-32) done = true;
        }
        if (trace) trace(x);
    }
    sum.final_cost = cost;
    return sum;
}

/* Ceres 1.14 LINE_SEARCH minimizer, BFGS direction, WOLFE line search, CUBIC interpolation, as
 * configured at AvatarOptimizer.cpp:1313-1341 (defaults listed in SURVEY Appendix B).  Restated from
 * the published algorithm (line_search_minimizer.cc / line_search.cc / line_search_direction.cc /
 * polynomial.cc); Ceres is not vendored: PARITY UNPINNED. */
struct FSample {
    double x = 0, value = 0, gradient = 0;
    bool value_is_valid = false, gradient_is_valid = false;
    std::vector<double> vector_x, vector_gradient;
};

// minimise the interpolating polynomial over [x_min, x_max] (polynomial.cc MinimizePolynomial)
double minimize_cubic_hermite(const FSample& s0, const FSample& s1, double x_min, double x_max) {
    // cubic a x^3 + b x^2 + c x + d through (x0,f0,g0), (x1,f1,g1)
    const double x0 = s0.x, x1 = s1.x, f0 = s0.value, f1 = s1.value, g0 = s0.gradient, g1 = s1.gradient;
    // solve the 4x4 Vandermonde-type system by Gaussian elimination with partial pivoting
    double M[4][5] = {{x0 * x0 * x0, x0 * x0, x0, 1, f0},
                      {3 * x0 * x0, 2 * x0, 1, 0, g0},
                      {x1 * x1 * x1, x1 * x1, x1, 1, f1},
                      {3 * x1 * x1, 2 * x1, 1, 0, g1}};
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
        for (int k = 0; k < 5; ++k) std::swap(M[c][k], M[piv][k]);
        if (M[c][c] == 0) continue;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = M[r][c] / M[c][c];
            for (int k = c; k < 5; ++k) M[r][k] -= f * M[c][k];
        }
    }
    double co[4];
    for (int c = 0; c < 4; ++c) co[c] = (M[c][c] != 0) ? M[c][4] / M[c][c] : 0.0;
    auto evalp = [&](double x) { return ((co[0] * x + co[1]) * x + co[2]) * x + co[3]; };
    double best_x = 0.5 * (x_min + x_max), best_v = evalp(best_x);
    auto consider = [&](double x) {
        const double v = evalp(x);
        if (v < best_v) {
            best_v = v;
            best_x = x;
        }
    };
    consider(x_min);
    consider(x_max);
    // roots of the derivative 3a x^2 + 2b x + c (real parts only, as Ceres does)
    const double a = 3 * co[0], b = 2 * co[1], c = co[2];
    if (a != 0) {
        const double disc = b * b - 4 * a * c;
        if (disc >= 0) {
            const double sq = std::sqrt(disc);
            const double r1 = (-b + sq) / (2 * a), r2 = (-b - sq) / (2 * a);
            if (r1 >= x_min && r1 <= x_max) consider(r1);
            if (r2 >= x_min && r2 <= x_max) consider(r2);
        } else {
            const double r = -b / (2 * a);
            if (r >= x_min && r <= x_max) consider(r);
        }
    } else if (b != 0) {
        const double r = -c / b;
        if (r >= x_min && r <= x_max) consider(r);
    }
    return best_x;
}

SolveSummary solve_bfgs_wolfe(int nx, int P, double* x, const EvalFn& eval, const PlusFn& plus, int max_iters,
                              double function_tolerance, const TraceFn& trace) {
    SolveSummary sum;
    const double sufficient_decrease = 1e-4, max_step_contraction = 1e-3, min_step_size = 1e-9;
    const double sufficient_curvature_decrease = 0.9, max_step_expansion = 10.0;
    const int max_ls_iters = 20, max_restarts = 5;
    const double gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    (void)max_step_contraction;

    struct State {
        double cost = 0, directional_derivative = 0, step_size = 0;
        std::vector<double> gradient, search_direction;
    };
    State cur, prev;
    cur.gradient.resize(P);
    cur.search_direction.assign(P, 0);
    cur.cost = eval(x, cur.gradient.data(), nullptr);
    sum.initial_cost = cur.cost;
    auto maxnorm = [](const std::vector<double>& v) {
        double m = 0;
        for (double e : v) m = std::max(m, std::fabs(e));
        return m;
    };
    std::vector<double> invH((size_t)P * P, 0.0);
    for (int i = 0; i < P; ++i) invH[(size_t)i * P + i] = 1.0;
    auto reset_bfgs = [&]() {
        std::fill(invH.begin(), invH.end(), 0.0);
        for (int i = 0; i < P; ++i) invH[(size_t)i * P + i] = 1.0;
    };
    std::vector<double> xd(nx), delta(P), b(P), dg(P), dx(P);
    int restarts = 0;
    bool ls_ok = true;
    if (maxnorm(cur.gradient) <= gradient_tolerance) {
        sum.final_cost = cur.cost;
        return sum;
    }
    // 1-D function phi(a) = f(x (+) a d)
    auto sample = [&](double a, const std::vector<double>& d, bool want_grad, FSample& s) {
        s.x = a;
        for (int i = 0; i < P; ++i) delta[i] = a * d[i];
        s.vector_x.resize(nx);
        plus(x, delta.data(), s.vector_x.data());
        s.vector_gradient.resize(P);
        s.value = eval(s.vector_x.data(), want_grad ? s.vector_gradient.data() : nullptr, nullptr);
        s.value_is_valid = std::isfinite(s.value);
        s.gradient_is_valid = false;
        if (want_grad && s.value_is_valid) {
            double gd = 0;
            for (int i = 0; i < P; ++i) gd += s.vector_gradient[i] * d[i];
            s.gradient = gd;
            s.gradient_is_valid = std::isfinite(gd);
        }
    };
    for (int iter = 1; iter <= max_iters; ++iter) {
        ++sum.iterations;
        bool dir_ok = true;
        if (iter == 1) {
            for (int i = 0; i < P; ++i) cur.search_direction[i] = -cur.gradient[i];
        } else {
            // BFGS::NextDirection
            for (int i = 0; i < P; ++i) {
                dx[i] = prev.search_direction[i] * prev.step_size;
                dg[i] = cur.gradient[i] - prev.gradient[i];
            }
            double dxdg = 0;
            for (int i = 0; i < P; ++i) dxdg += dx[i] * dg[i];
            if (dxdg > 1e-14) {
                const double rho = 1.0 / dxdg;
                double dgb = 0;
                for (int i = 0; i < P; ++i) {
                    double s = 0;
                    for (int j = 0; j < P; ++j) s += invH[(size_t)i * P + j] * dg[j];
                    b[i] = s;
                }
                for (int i = 0; i < P; ++i) dgb += dg[i] * b[i];
                const double c = (dxdg + dgb) * rho * rho;
                for (int i = 0; i < P; ++i)
                    for (int j = 0; j < P; ++j)
                        invH[(size_t)i * P + j] += c * dx[i] * dx[j] - rho * (b[i] * dx[j] + dx[i] * b[j]);
            }
            double dd = 0;
            for (int i = 0; i < P; ++i) {
                double s = 0;
                for (int j = 0; j < P; ++j) s -= invH[(size_t)i * P + j] * cur.gradient[j];
                cur.search_direction[i] = s;
                dd += s * cur.gradient[i];
            }
            dir_ok = dd < 0;
            if (!dir_ok) {
                if (restarts >= max_restarts) break;
                ++restarts;
                reset_bfgs();
                for (int i = 0; i < P; ++i) cur.search_direction[i] = -cur.gradient[i];
            } else {
                restarts = 0;
            }
        }
        cur.directional_derivative = 0;
        for (int i = 0; i < P; ++i) cur.directional_derivative += cur.gradient[i] * cur.search_direction[i];
        const double gmx = maxnorm(cur.gradient);
        double step0 = (iter == 1 || !ls_ok || !dir_ok)
                           ? std::min(1.0, 1.0 / gmx)
                           : std::min(1.0, 2.0 * (cur.cost - prev.cost) / cur.directional_derivative);
        if (!(step0 >= min_step_size)) break;
        // ---- WolfeLineSearch::DoSearch ----
        FSample initial;
        initial.x = 0;
        initial.value = cur.cost;
        initial.gradient = cur.directional_derivative;
        initial.value_is_valid = initial.gradient_is_valid = true;
        FSample low, high, current, previous = initial, solution;
        bool do_zoom = false, bracket_ok = false;
        int ls_iter = 0;
        sample(step0, cur.search_direction, true, current);
        while (true) {  // BracketingPhase
            ++ls_iter;
            if (current.value_is_valid &&
                (current.value > initial.value + sufficient_decrease * initial.gradient * current.x ||
                 (previous.value_is_valid && current.value > previous.value))) {
                do_zoom = true; low = previous; high = current; bracket_ok = true;
                break;
            }
            if (current.value_is_valid && std::fabs(current.gradient) <= -sufficient_curvature_decrease * initial.gradient) {
                low = current; high = current; do_zoom = false; bracket_ok = true;
                break;
            }
            if (current.value_is_valid && current.gradient >= 0) {
                do_zoom = true; low = current; high = previous; bracket_ok = true;
                break;
            }
            if (ls_iter >= max_ls_iters) {
                low = (current.value_is_valid && current.value < previous.value) ? current : previous;
                bracket_ok = low.x > 0;
                do_zoom = false;
                break;
            }
            double next;
            if (current.value_is_valid) {
                next = minimize_cubic_hermite(previous, current, current.x, current.x * max_step_expansion);
                if (!(next > current.x)) next = current.x * max_step_expansion;
                previous = current;
            } else {
                next = current.x * 0.5;
                if (next < min_step_size) break;
            }
            sample(next, cur.search_direction, true, current);
        }
        FSample optimal;
        ls_ok = bracket_ok;
        if (bracket_ok && !do_zoom) {
            optimal = low;
        } else if (bracket_ok) {  // ZoomPhase
            bool found = false;
            while (ls_iter < max_ls_iters) {
                ++ls_iter;
                if (std::fabs(high.x - low.x) < min_step_size) break;
                const double lo = std::min(low.x, high.x), hi = std::max(low.x, high.x);
                double a = (low.gradient_is_valid && high.gradient_is_valid)
                               ? minimize_cubic_hermite(low, high, lo, hi)
                               : 0.5 * (lo + hi);
                sample(a, cur.search_direction, true, solution);
                if (!solution.value_is_valid) break;
                if (solution.value > initial.value + sufficient_decrease * initial.gradient * solution.x ||
                    solution.value >= low.value) {
                    high = solution;
                    continue;
                }
                if (std::fabs(solution.gradient) <= -sufficient_curvature_decrease * initial.gradient) {
                    found = true;
                    break;
                }
                if (solution.gradient * (high.x - low.x) >= 0) high = low;
                low = solution;
            }
            optimal = (found && solution.value_is_valid && solution.value <= low.value) ? solution : low;
            if (!(optimal.x > 0)) ls_ok = false;
        }
        if (!ls_ok) break;  // "line search failed" terminates the minimizer
        // accept: Ceres 1.14 re-evaluates cost and gradient at x_plus_delta
        prev = cur;
        prev.step_size = optimal.x;
        for (int i = 0; i < P; ++i) delta[i] = optimal.x * cur.search_direction[i];
        plus(x, delta.data(), xd.data());
        double dn = 0, xn = 0;
        for (int i = 0; i < nx; ++i) {
            dn += (xd[i] - x[i]) * (xd[i] - x[i]);
            xn += x[i] * x[i];
        }
        std::copy(xd.begin(), xd.end(), x);
        cur.cost = eval(x, cur.gradient.data(), nullptr);
        ++sum.accepted;
        if (trace) trace(x);
        if (maxnorm(cur.gradient) <= gradient_tolerance) break;
        if (std::fabs(prev.cost - cur.cost) <= function_tolerance * prev.cost) break;
        if (std::sqrt(dn) <= parameter_tolerance * (std::sqrt(xn) + parameter_tolerance)) break;
    }
    sum.final_cost = cur.cost;
    return sum;
}

}  // namespace

/* ============================ C API ============================ */
extern "C" {

orc_model* orc_model_create(int V, int J, int K, int F, const double* v_template, const double* shapedirs,
                            const double* j_regressor, const double* weights, const int32_t* parents,
                            const int32_t* faces) {
    auto* m = new orc_model;
    m->V = V; m->J = J; m->K = K; m->F = F;
    m->parent.assign(parents, parents + J);
    m->baseCloud.assign(v_template, v_template + 3 * (size_t)V);         // AvatarModel.cpp:44-47
    m->mesh.assign(faces, faces + 3 * (size_t)F);                          // :50-53
    m->keyClouds.assign(shapedirs, shapedirs + 3 * (size_t)V * K);         // :97-103
    m->assignedJoints.resize(V);                                           // :74-94
    m->weightsCol.resize(V);
    for (int v = 0; v < V; ++v)
        for (int j = 0; j < J; ++j) {
            const double wt = weights[(size_t)v * J + j];
            if (wt != 0.0) m->weightsCol[v].push_back({j, wt});  // sparseView drops exact zeros
            if (wt != 0.0 && wt > 1e-12) m->assignedJoints[v].push_back({wt, j});
        }
    for (int v = 0; v < V; ++v)
        std::sort(m->assignedJoints[v].begin(), m->assignedJoints[v].end(), std::greater<std::pair<double, int>>());
    // joint shape regressor (:111-127): jointRegressor is sparse (V x J); products sum over its nonzeros
    m->initialJointPos.assign(3 * (size_t)J, 0);
    m->jointShapeReg.assign(3 * (size_t)J * K, 0);
    for (int j = 0; j < J; ++j)
        for (int v = 0; v < V; ++v) {
            const double r = j_regressor[(size_t)j * V + v];
            if (r == 0.0) continue;
            for (int c = 0; c < 3; ++c) {
                m->initialJointPos[3 * j + c] += m->baseCloud[3 * (size_t)v + c] * r;
                for (int k = 0; k < K; ++k)
                    m->jointShapeReg[((size_t)3 * j + c) * K + k] += m->keyClouds[((size_t)3 * v + c) * K + k] * r;
            }
        }
    m->jointShapeRegBase = m->initialJointPos;
    return m;
}
void orc_model_destroy(orc_model* m) { delete m; }

int orc_model_set_prior(orc_model* m, int C, int D, const double* weight, const double* mean, const double* cov) {
    GMM& g = m->posePrior;
    g.nComps = C; g.nDims = D;
    g.weight.assign(weight, weight + C);
    g.mean.assign(mean, mean + (size_t)C * D);
    g.cov.assign(cov, cov + (size_t)C * D * D);
    return g.finish_load();
}
int orc_model_load_prior_text(orc_model* m, const char* path) { return m->posePrior.load_text(path); }

void orc_model_get_joint_reg(const orc_model* m, double* base, double* reg, double* initial) {
    if (base) std::copy(m->jointShapeRegBase.begin(), m->jointShapeRegBase.end(), base);
    if (reg) std::copy(m->jointShapeReg.begin(), m->jointShapeReg.end(), reg);
    if (initial) std::copy(m->initialJointPos.begin(), m->initialJointPos.end(), initial);
}
int orc_model_get_assigned(const orc_model* m, int32_t* start, int32_t* joint, double* weight) {
    int n = 0;
    for (int v = 0; v < m->V; ++v) {
        if (start) start[v] = n;
        for (auto& wj : m->assignedJoints[v]) {
            if (joint) joint[n] = wj.second;
            if (weight) weight[n] = wj.first;
            ++n;
        }
    }
    if (start) start[m->V] = n;
    return n;
}
void orc_model_get_prior(const orc_model* m, double* prec_cho, double* consts_log) {
    const GMM& g = m->posePrior;
    if (prec_cho) std::copy(g.prec_cho.begin(), g.prec_cho.end(), prec_cho);
    if (consts_log) std::copy(g.consts_log.begin(), g.consts_log.end(), consts_log);
}

void orc_avatar_update(const orc_model* m, const double* p, const double* R, const double* w, double* cloud,
                       double* joint_pos, double* joint_trans) {
    std::vector<M3> Rm(m->J);
    for (int j = 0; j < m->J; ++j) std::copy(R + 9 * j, R + 9 * j + 9, Rm[j].m);
    avatar_update(*m, p, Rm.data(), w, cloud, joint_pos, joint_trans);
}
void orc_rotmat_to_quat(const double* R, double* q) {
    M3 m;
    std::copy(R, R + 9, m.m);
    double ang; V3 ax;
    quat_to_angle_axis(rot_to_quat_raw(m), ang, ax);  // AngleAxisd::fromRotationMatrix
    const Quat r = angle_axis_to_quat(ang, ax);       // Quaterniond = AngleAxisd
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
void orc_quat_to_rotmat(const double* q, double* R) {
    const M3 m = quat_to_rot(Quat{q[0], q[1], q[2], q[3]});
    std::copy(m.m, m.m + 9, R);
}
int orc_gmm_residual(const orc_model* m, const double* x, double* out) { return m->posePrior.residual(x, out); }

orc_optimizer* orc_optimizer_create(const orc_model* m, int num_parts, const int32_t* part_map) {
    auto* o = new orc_optimizer;
    o->model = m;
    o->numParts = num_parts;
    o->partMap.assign(part_map, part_map + m->J);
    o->modelPartIndices.resize(num_parts);
    o->modelPartLabels.resize(m->V);
    for (int i = 0; i < m->V; ++i) {  // :1227-1243, :1307-1311
        const int mainJointId = m->assignedJoints[i][0].second;
        const int partId = o->partMap[mainJointId];
        o->modelPartIndices[partId].push_back(i);
        o->modelPartLabels[i] = partId;
    }
    return o;
}
void orc_optimizer_destroy(orc_optimizer* o) { delete o; }
int orc_param_dim(const orc_optimizer* o) { return 3 + 4 * o->model->J + o->model->K; }
int orc_tangent_dim(const orc_optimizer* o) { return 3 + 3 * o->model->J + o->model->K; }

void orc_visibility(const orc_optimizer* o, const double* cloud, uint8_t* vis) { visibility(*o->model, cloud, vis); }

void orc_find_nn(const orc_optimizer* o, const double* cloud, const uint8_t* vis, const double* data,
                 const int32_t* labels, int N, int method, int32_t* idx) {
    find_nn(*o, cloud, vis, data, labels, N, method, 1, idx);
}

/* plain exact 1-NN of every query among pts[0..n) with the distance arithmetic of nanoflann.hpp:423-445 (strict <,
 * lowest index wins exact ties); parallel over queries */
void orc_brute_nn(const double* pts, int n, const double* queries, int nq, int32_t* out) {
    const int nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() {
            for (int i = t; i < nq; i += nt) {
                double best = std::numeric_limits<double>::max();
                int bi = -1;
                for (int k = 0; k < n; ++k) {
                    const double d = dist2(queries + 3 * (size_t)i, pts + 3 * (size_t)k);
                    if (d < best) {
                        best = d;
                        bi = k;
                    }
                }
                out[i] = bi;
            }
        });
    for (auto& x : th) x.join();
}

static void build_problem(Common& cm, Problem& pb, const int32_t* idx, int N) {
    const int V = cm.V;
    std::vector<std::vector<int>> correspondences(V);
    for (int i = 0; i < N; ++i)
        if (idx[i] >= 0) correspondences[idx[i]].push_back(i);
    cm.caches.clear();
    pb.corr.clear();
    pb.totalResiduals = 0;
    for (int i = 0; i < V; ++i) {  // :1419-1451
        if (correspondences[i].empty()) continue;
        cm.make_cache(i);
        pb.totalResiduals += correspondences[i].size();
        pb.corr.push_back(std::move(correspondences[i]));
    }
    cm.scaledBetaPose = pb.betaPose * std::sqrt((double)pb.totalResiduals) / 15.;    // :1457
    cm.scaledBetaShape = pb.betaShape * std::sqrt((double)pb.totalResiduals) / 15.;  // :1458
}

double orc_evaluate(const orc_optimizer* o, const double* x, const double* data, const int32_t* idx, int N,
                    double beta_pose, double beta_shape, int num_threads, double* grad, double* H) {
    Common cm(*o->model);
    cm.numThreads = num_threads;
    Problem pb(cm, data, beta_pose, beta_shape);
    build_problem(cm, pb, idx, N);
    return pb.evaluate(x, grad, H);
}

void orc_vertex_jacobian(const orc_optimizer* o, const double* x, int vertex, double* pos3, double* jac) {
    Common cm(*o->model);
    const int J = cm.J, K = cm.K, P = 3 + 3 * J + K;
    cm.make_cache(vertex);
    cm.set_params(x);
    cm.PrepareForEvaluation(true);
    const Cache& ch = cm.caches[0];
    for (int c = 0; c < 3; ++c) pos3[c] = ch.resid[c];
    if (!jac) return;
    std::fill(jac, jac + 3 * (size_t)P, 0.0);
    const auto& anc = cm.ancestor[vertex];
    for (int r = 0; r < 3; ++r) {
        jac[(size_t)r * P + r] = 1.0;
        for (size_t a = 0; a < anc.size(); ++a)
            for (int c = 0; c < 3; ++c) jac[(size_t)r * P + 3 + 3 * anc[a].jid + c] = ch.icpJacobian[a](r, c);
        for (int k = 0; k < K; ++k) jac[(size_t)r * P + 3 + 3 * J + k] = ch.icpShapeJacobian[r * K + k];
    }
}

// AvatarICPAutoDiffCostFunctor::operator() (AvatarOptimizer.cpp:742-818), without the data term
void orc_vertex_position_chain(const orc_optimizer* o, const double* x, int vertex, double* pos3) {
    const orc_model& md = *o->model;
    const int J = md.J, K = md.K;
    const double* w = x + 3 + 4 * J;
    auto shaped_point = [&](int v, int c) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += md.keyClouds[((size_t)3 * v + c) * K + k] * w[k];
        return s + md.baseCloud[3 * (size_t)v + c];
    };
    std::vector<double> jointPos(3 * (size_t)J);
    for (int r = 0; r < 3 * J; ++r) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += md.jointShapeReg[(size_t)r * K + k] * w[k];
        jointPos[r] = md.jointShapeRegBase[r] + s;
    }
    const V3 offset{jointPos[0], jointPos[1], jointPos[2]};
    V3 resid{0, 0, 0};
    for (auto& assign : md.assignedJoints[vertex]) {
        const int k0 = assign.second;
        V3 vec{shaped_point(vertex, 0) - offset[0] - (jointPos[3 * k0] - offset[0]),
               shaped_point(vertex, 1) - offset[1] - (jointPos[3 * k0 + 1] - offset[1]),
               shaped_point(vertex, 2) - offset[2] - (jointPos[3 * k0 + 2] - offset[2])};
        for (int p = k0; p != -1; p = md.parent[p]) {
            const M3 Rq = quat_to_rot(Quat{x[3 + 4 * p], x[4 + 4 * p], x[5 + 4 * p], x[6 + 4 * p]});
            vec = mul(Rq, vec);
            if (p) {
                const int pp = md.parent[p];
                for (int c = 0; c < 3; ++c) vec[c] += jointPos[3 * p + c] - jointPos[3 * pp + c];
            }
        }
        resid = add(resid, scale(vec, assign.first));
    }
    for (int c = 0; c < 3; ++c) pos3[c] = resid[c] + x[c];
}

void orc_retract(const orc_optimizer* o, const double* x, const double* delta, double* xp) {
    Common cm(*o->model);
    Problem pb(cm, nullptr, 0, 0);
    pb.plus(x, delta, xp);
}

int orc_optimize(const orc_optimizer* o, const double* data, const int32_t* labels, int N, double* x,
                 const orc_options* op, orc_stats* st, double* trace, int trace_cap, int* trace_len,
                 int32_t* nn_out) {
    const auto t0 = std::chrono::steady_clock::now();
    const orc_model& md = *o->model;
    const int V = md.V, J = md.J, K = md.K;
    const int nx = 3 + 4 * J + K, P = 3 + 3 * J + K;
    Common cm(md);  // :1256 (rebuilt every optimize() call, like the reference)
    cm.numThreads = op->num_threads;
    std::vector<uint8_t> vis(V, 1);
    std::vector<double> cloud(3 * (size_t)V);
    std::vector<M3> Rm(J);
    std::vector<int32_t> idx(N);
    int ntrace = 0;
    long evals = 0;
    SolveSummary ss;
    size_t ncorr = 0;
    auto refresh_cloud = [&]() {  // :1494-1497
        for (int j = 0; j < J; ++j) Rm[j] = quat_to_rot(Quat{x[3 + 4 * j], x[4 + 4 * j], x[5 + 4 * j], x[6 + 4 * j]});
        avatar_update(md, x, Rm.data(), x + 3 + 4 * J, cloud.data(), nullptr, nullptr);
    };
    refresh_cloud();  // the caller's ava.update()
    for (int icp = 0; icp < op->icp_iters; ++icp) {
        if (op->enable_occlusion) visibility(md, cloud.data(), vis.data());  // :1349-1367
        find_nn(*o, cloud.data(), vis.data(), data, labels, N, op->nn_method, op->num_threads, idx.data());
        Problem pb(cm, data, op->beta_pose, op->beta_shape);
        build_problem(cm, pb, idx.data(), N);
        ncorr = pb.totalResiduals;
        EvalFn ev = [&](const double* xx, double* g, double* H) { return pb.evaluate(xx, g, H); };
        PlusFn pl = [&](const double* xx, const double* d, double* xp) { pb.plus(xx, d, xp); };
        TraceFn tr = [&](const double* xx) {
            if (trace && ntrace < trace_cap) std::copy(xx, xx + nx, trace + (size_t)nx * ntrace);
            ++ntrace;
        };
        if (op->solver == ORC_SOLVER_GN_LM)
            ss = solve_gn_lm(nx, P, x, ev, pl, op->max_iters_per_icp, op->function_tolerance, tr);
        else
            ss = solve_bfgs_wolfe(nx, P, x, ev, pl, op->max_iters_per_icp, op->function_tolerance, tr);
        evals += pb.evaluations;
        refresh_cloud();
    }
    if (nn_out) std::copy(idx.begin(), idx.end(), nn_out);
    if (trace_len) *trace_len = std::min(ntrace, trace_cap);
    if (st) {
        st->num_correspondences = (int)ncorr;
        st->iterations = ss.iterations;
        st->accepted_steps = ss.accepted;
        st->evaluations = (int)evals;
        st->initial_cost = ss.initial_cost;
        st->final_cost = ss.final_cost;
        st->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return 0;
}


/* demo.cpp:215-250 + Calibration.cpp:83-95 (see header) */
int64_t orc_build_cloud(const float* depth, const uint8_t* parts, int width, int height, const float* intrin,
                        const int32_t* roi, int interval, int num_parts, double* cloud, int32_t* labels,
                        int64_t capacity) {
    const float fx = intrin[0], cx = intrin[1], fy = intrin[2], cy = intrin[3];
    int x0 = 0, y0 = 0, x1 = width - 1, y1 = height - 1;
    if (roi) { x0 = roi[0]; y0 = roi[1]; x1 = roi[2]; y1 = roi[3]; }
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > width - 1) x1 = width - 1;
    if (y1 > height - 1) y1 = height - 1;
    // first pass: count (demo.cpp:216-225)
    int64_t cnz = 0;
    for (int r = y0; r <= y1; r += interval) {
        const uint8_t* partptr = parts + (size_t)r * width;
        for (int c = x0; c <= x1; c += interval) {
            if (partptr[c] == 255) continue;
            ++cnz;
        }
    }
    if (cnz > capacity) return -2;
    // second pass: fill (demo.cpp:229-250); xyz as depthToXYZ computes it (Calibration.cpp:91)
    int64_t i = 0;
    for (int r = y0; r <= y1; r += interval) {
        const float* inPtr = depth + (size_t)r * width;
        const uint8_t* partptr = parts + (size_t)r * width;
        for (int c = x0; c <= x1; c += interval) {
            if (partptr[c] == 255) continue;
            if (partptr[c] >= num_parts) return -1;
            const float z = inPtr[c];
            const float X = (c - cx) * z / fx;
            const float Y = (r - cy) * z / fy;
            cloud[3 * i] = X;
            cloud[3 * i + 1] = -Y;
            cloud[3 * i + 2] = z;
            labels[i] = partptr[c];
            ++i;
        }
    }
    return cnz;
}

/* RTree::postProcess (RTree.cpp:3422-3450) with suppressPartNonMax (:125-232) / removeSmallPieces (:234-320) and the
 * trailing upscaleGrid (:70-100), restated literally -- INCLUDING the reference's quirk for interval > 1: the downward
 * probe tests pixel (r + 1, c) but pushes the id of (r + interval, c), unmarked and whatever its label (:176), so a
 * component "passes through" the grid pixel below; such pixels are counted, can be erased with the component and can
 * still seed their own component later.  The scan is sequential and its result depends on the visiting order, which is
 * therefore kept: raster order of the seeds, LIFO stack, probes in the order up, down, left, right.
 * image [H][W] uint8 in/out (255 = background); roi = {x0, y0, x1, y1} inclusive or NULL; com_pre [2 * num_parts] in/out,
 * column major like the reference's 2 x numParts matrix (x of part i at [2 i], y at [2 i + 1]; x = -1: part not seen);
 * part_map_type 0 = contiguous (keep the best blob per part), otherwise remove pieces smaller than 0.0005 of the grid. */
void orc_rtree_postprocess(uint8_t* image, int width, int height, const int32_t* roi, int interval, int num_parts, int part_map_type,
                           double* com_pre, double dist_to_pre_weight) {
    int tlx = 0, tly = 0, brx = width - 1, bry = height - 1;
    if (roi) { tlx = roi[0]; tly = roi[1]; brx = roi[2]; bry = roi[3]; }
    auto at = [&](int r, int c) -> uint8_t& { return image[(size_t)r * width + c]; };
    const int VISITED_OFFSET = 128;
    std::vector<int> stk, curCompVis;
    int hi_bit = (1 << 16);
    const int lo_mask = hi_bit - 1;
    hi_bit *= interval;
    uint8_t seed_val = 0;
    auto maybe_visit = [&](int new_r, int new_c, int new_id) {
        uint8_t& val = at(new_r, new_c);
        if (seed_val == val) {
            val += VISITED_OFFSET;
            curCompVis.push_back(new_id);
            stk.push_back(new_id);
        }
    };
    if (part_map_type == 0) {
        std::vector<std::vector<int>> bestComp(num_parts);
        std::vector<double> bestScore(num_parts, 0.0), comBest(2 * (size_t)num_parts, 0.0);
        for (int rr = tly; rr <= bry; rr += interval) {
            for (int cc = tlx; cc <= brx; cc += interval) {
                const uint8_t val = at(rr, cc);
                if (val >= VISITED_OFFSET) continue;
                seed_val = val;
                at(rr, cc) += VISITED_OFFSET;
                stk.push_back((rr << 16) + cc);
                curCompVis.clear();
                curCompVis.push_back(stk.back());
                double com0 = 0, com1 = 0;
                const bool hasPrevCom = com_pre[2 * val] >= 0.;
                while (stk.size()) {
                    const int id = stk.back();
                    const int cur_c = (id & lo_mask), cur_r = (id >> 16);
                    stk.pop_back();
                    if (cur_r >= tly + interval) maybe_visit(cur_r - interval, cur_c, id - hi_bit);
                    if (cur_r <= bry - interval) maybe_visit(cur_r + 1, cur_c, id + hi_bit);
                    if (cur_c >= tlx + interval) maybe_visit(cur_r, cur_c - interval, id - interval);
                    if (cur_c <= brx - interval) maybe_visit(cur_r, cur_c + interval, id + interval);
                    com0 += cur_c;
                    com1 += cur_r;
                }
                double score = (double)curCompVis.size();
                com0 /= curCompVis.size();
                com1 /= curCompVis.size();
                if (hasPrevCom) {
                    const double d0 = com0 - com_pre[2 * val], d1 = com1 - com_pre[2 * val + 1];
                    score -= (d0 * d0 + d1 * d1) * dist_to_pre_weight;
                }
                if (score > bestScore[val]) {
                    bestScore[val] = score;
                    comBest[2 * val] = com0;
                    comBest[2 * val + 1] = com1;
                    for (int id : bestComp[val]) at(id >> 16, id & lo_mask) = 255;
                    bestComp[val].swap(curCompVis);
                } else {
                    for (int id : curCompVis) at(id >> 16, id & lo_mask) = 255;
                }
            }
        }
        for (int i = 0; i < num_parts; ++i) {
            if (bestComp[i].empty()) {
                com_pre[2 * i] = -1.;
            } else {
                com_pre[2 * i] = comBest[2 * i];
                com_pre[2 * i + 1] = comBest[2 * i + 1];
            }
        }
    } else {
        const size_t scaledThresh = (size_t)(height * width / (interval * interval) * 0.0005);
        for (int rr = tly; rr <= bry; rr += interval) {
            for (int cc = tlx; cc <= brx; cc += interval) {
                const uint8_t val = at(rr, cc);
                if (val >= VISITED_OFFSET) continue;
                seed_val = val;
                at(rr, cc) += VISITED_OFFSET;
                stk.push_back((rr << 16) + cc);
                curCompVis.clear();
                curCompVis.push_back(stk.back());
                while (stk.size()) {
                    const int id = stk.back();
                    const int cur_c = (id & lo_mask), cur_r = (id >> 16);
                    stk.pop_back();
                    if (cur_r >= tly + interval) maybe_visit(cur_r - interval, cur_c, id - hi_bit);
                    if (cur_r <= bry - interval) maybe_visit(cur_r + 1, cur_c, id + hi_bit);
                    if (cur_c >= tlx + interval) maybe_visit(cur_r, cur_c - interval, id - interval);
                    if (cur_c <= brx - interval) maybe_visit(cur_r, cur_c + interval, id + interval);
                }
                if (curCompVis.size() < scaledThresh)
                    for (int id : curCompVis) at(id >> 16, id & lo_mask) = 255;
            }
        }
    }
    for (int r = tly; r <= bry; ++r)
        for (int cc = tlx; cc <= brx; ++cc) {
            uint8_t& val = at(r, cc);
            if (val >= VISITED_OFFSET && val != 255) val -= VISITED_OFFSET;
        }
    if (interval > 1) {   // upscaleGrid (RTree.cpp:70-100); memset(ptr + cc, val, interval) may pass bot_right.x: clamped to the image
        for (int rr = tly + interval; rr <= bry; rr += interval) {
            const uint8_t* ptrRef = image + (size_t)rr * width;
            for (int r = rr; r < rr + interval; ++r) {
                if (r > bry) break;
                uint8_t* ptr = image + (size_t)r * width;
                for (int cc = tlx; cc <= brx; cc += interval) {
                    const uint8_t v = ptrRef[cc];
                    for (int k = 0; k < interval && cc + k < width; ++k) ptr[cc + k] = v;
                }
            }
        }
    }
}

/* RTree::predictBest on an image + upscaleGrid (see header) */
void orc_rtree_predict(const float* depth, int width, int height, int num_nodes, const float* u, const float* v,
                       const float* thresh, const int32_t* lnode, const int32_t* rnode, const int32_t* leafid,
                       const uint8_t* leaf_best, const int32_t* roi, int interval, int fill_in_gaps, uint8_t* out) {
    (void)num_nodes;
    const float BACKGROUND_DEPTH = 20.f;   // RTree.cpp:325
    std::fill(out, out + (size_t)width * height, (uint8_t)255);   // result.setTo(255), RTree.cpp:3189
    int tlx = 0, tly = 0, brx = width - 1, bry = height - 1;      // bot_right == -1: whole image (:3190-3193)
    if (roi) { tlx = roi[0]; tly = roi[1]; brx = roi[2]; bry = roi[3]; }
    // RTree.cpp:3196-3199: r = (row += interval) with row starting at top_left.y
    for (int r = tly + interval; r <= bry; r += interval) {
        if (r < 0 || r >= height) continue;
        const float* inPtr = depth + (size_t)r * width;
        for (int c = tlx; c <= brx; c += interval) {
            if (c < 0 || c >= width) continue;
            if (inPtr[c] == 0.f) continue;
            int nodeid = 0;
            const float sampleDepth = inPtr[c];
            while (leafid[nodeid] == -1) {
                // Eigen Vector2f / float: component-wise division; std::round: half away from zero (:3209-3217)
                const float utx = u[2 * nodeid] / sampleDepth, uty = u[2 * nodeid + 1] / sampleDepth;
                const float vtx = v[2 * nodeid] / sampleDepth, vty = v[2 * nodeid + 1] / sampleDepth;
                const int32_t ux = static_cast<int32_t>(std::round(utx)) + c, uy = static_cast<int32_t>(std::round(uty)) + r;
                const int32_t vx = static_cast<int32_t>(std::round(vtx)) + c, vy = static_cast<int32_t>(std::round(vty)) + r;
                float zu, zv;
                if (ux < tlx || uy < tly || ux > brx || uy > bry) {
                    zu = BACKGROUND_DEPTH;
                } else {
                    zu = depth[(size_t)uy * width + ux];
                    if (zu == 0.0) zu = BACKGROUND_DEPTH;
                }
                if (vx < tlx || vy < tly || vx > brx || vy > bry) {
                    zv = BACKGROUND_DEPTH;
                } else {
                    zv = depth[(size_t)vy * width + vx];
                    if (zv == 0.0) zv = BACKGROUND_DEPTH;
                }
                nodeid = (zu - zv < thresh[nodeid]) ? lnode[nodeid] : rnode[nodeid];
            }
            out[(size_t)r * width + c] = leaf_best[leafid[nodeid]];
        }
    }
    if (fill_in_gaps && interval > 1) {   // upscaleGrid, RTree.cpp:70-100
        for (int rr = tly + interval; rr <= bry; rr += interval) {
            if (rr < 0 || rr >= height) continue;
            const uint8_t* ptrRef = out + (size_t)rr * width;
            for (int r = rr; r < rr + interval; ++r) {
                if (r > bry) break;
                if (r >= height) break;
                uint8_t* ptr = out + (size_t)r * width;
                for (int cc = tlx; cc <= brx; cc += interval) {
                    if (cc < 0 || cc >= width) continue;
                    const uint8_t val = ptrRef[cc];
                    for (int x = cc; x < cc + interval && x < width; ++x) ptr[x] = val;   // memset(ptr + cc, val, interval)
                }
            }
        }
    }
}
}  // extern "C"
