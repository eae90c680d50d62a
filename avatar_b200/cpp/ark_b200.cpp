// ark_b200.cpp -- implementation of the header-compatible C++ facade (include/ark_b200/*.h) over the C ABI.
// Host-side only: file parsing and marshalling; every computation happens in libavatar_b200.so.
#include "../../include/ark_b200/AvatarOptimizer.h"
#include "../../include/ark_b200/npz.h"
#include "../../include/ark_b200/RTree.h"
#include "../../include/ark_b200/AvatarRenderer.h"
#include "../../include/avatar_b200.h"
#include <random>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>

namespace ark {
namespace {
[[noreturn]] void die(const std::string& what) { throw std::runtime_error("avatar_b200: " + what + ": " + avb_last_error()); }
}

// GaussianMixture.cpp:12-58 (parsing only)
void GaussianMixture::load(const std::string& path) {
    FILE* fp = std::fopen(path.c_str(), "r");
    if (!fp) {
        std::fprintf(stderr, "Warning: pose prior file at %s does not exist or cannot be read\n", path.c_str());
        nComps = -1;
        return;
    }
    bool ok = std::fscanf(fp, "%d %d", &nComps, &nDims) == 2;
    weight.resize(nComps);
    mean.resize(nComps, nDims);
    cov.assign(nComps, Eigen::MatrixXd());
    for (int i = 0; i < nComps && ok; ++i) ok = std::fscanf(fp, "%lf", &weight(i)) == 1;
    for (int i = 0; i < nComps && ok; ++i)
        for (int j = 0; j < nDims && ok; ++j) ok = std::fscanf(fp, "%lf", &mean(i, j)) == 1;
    for (int i = 0; i < nComps && ok; ++i) {
        cov[i].resize(nDims, nDims);
        for (int j = 0; j < nDims && ok; ++j)
            for (int k = 0; k < nDims && ok; ++k) ok = std::fscanf(fp, "%lf", &cov[i](j, k)) == 1;
    }
    std::fclose(fp);
    if (!ok) throw std::runtime_error("pose prior file is truncated: " + path);
}

// AvatarModel.cpp:14-127 (npz branch) + :296
AvatarModel::AvatarModel(const std::string& model_dir, bool /*limit_one_joint_per_point: ignored on the npz path*/)
    : MODEL_DIR(model_dir) {
    std::string root = model_dir;
    if (root.empty()) {
        const char* env = std::getenv("OPENARK_DIR");  // Util.cpp:70-96
        root = std::string(env ? env : ".") + "/data/avatar-model";
    }
    auto npz = ark_b200::load_npz(root + "/model.npz");
    const auto& vt = npz.at("v_template");
    const auto& kt = npz.at("kintree_table");
    const auto& fa = npz.at("f");
    const auto& sd = npz.at("shapedirs");
    const auto& jr = npz.at("J_regressor");
    const auto& wt = npz.at("weights");
    const size_t V = vt.shape[0], J = kt.shape[1], F = fa.shape[0], K = sd.shape[2];
    parent.resize((long)J);
    for (size_t j = 0; j < J; ++j) parent((long)j) = (int)kt.as<uint32_t>(j);  // cast<int>(): 0xFFFFFFFF -> -1
    if (parent(0) != -1) throw std::runtime_error("model.npz: parent[0] must be -1");
    baseCloud.resize((long)(3 * V));
    for (size_t i = 0; i < 3 * V; ++i) baseCloud((long)i) = vt.as<double>(i);
    mesh.resize(3, (long)F);
    for (size_t f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c) mesh(c, (long)f) = (int)fa.as<int64_t>(3 * f + c);
    keyClouds.resize((long)(3 * V), (long)K);
    for (size_t r = 0; r < 3 * V; ++r)
        for (size_t k = 0; k < K; ++k) keyClouds((long)r, (long)k) = sd.as<double>(r * K + k);
    assignedJoints.assign(V, {});
    assignedPoints.assign(J, {});
    for (size_t v = 0; v < V; ++v)
        for (size_t j = 0; j < J; ++j) {
            const double w = wt.as<double>(v * J + j);
            if (w > 1e-12) {  // AvatarModel.cpp:81
                assignedJoints[v].push_back({w, (int)j});
                assignedPoints[j].push_back({w, (int)v});
            }
        }
    for (auto& a : assignedPoints) std::sort(a.begin(), a.end(), std::greater<std::pair<double, int>>());
    for (auto& a : assignedJoints) std::sort(a.begin(), a.end(), std::greater<std::pair<double, int>>());
    // joint shape regressor (AvatarModel.cpp:105-127)
    useJointShapeRegressor = true;
    initialJointPos.resize(3, (long)J);
    jointShapeReg.resize((long)(3 * J), (long)K);
    initialJointPos.setZero();
    jointShapeReg.setZero();
    for (size_t j = 0; j < J; ++j)
        for (size_t v = 0; v < V; ++v) {
            const double rg = jr.as<double>(j * V + v);
            if (rg == 0.0) continue;
            for (int c = 0; c < 3; ++c) {
                initialJointPos(c, (long)j) += baseCloud((long)(3 * v + c)) * rg;
                for (size_t k = 0; k < K; ++k) jointShapeReg((long)(3 * j + c), (long)k) += keyClouds((long)(3 * v + c), (long)k) * rg;
            }
        }
    jointShapeRegBase.resize((long)(3 * J));
    for (size_t j = 0; j < J; ++j)
        for (int c = 0; c < 3; ++c) jointShapeRegBase((long)(3 * j + c)) = initialJointPos(c, (long)j);
    posePrior.load(root + "/pose_prior.txt");
}

AvatarModel::~AvatarModel() {
    if (handle_) avb_model_destroy(handle_);
}

avb_model* AvatarModel::handle() const {
    if (handle_) return handle_;
    const int V = numPoints(), J = numJoints(), K = numShapeKeys(), F = numFaces();
    std::vector<int32_t> start{0}, joint, faces(3 * (size_t)F), par(J);
    std::vector<double> weight, key((size_t)3 * V * K), jsr((size_t)3 * J * K), mean, cov;
    for (auto& pairs : assignedJoints) {
        for (auto& wj : pairs) {
            weight.push_back(wj.first);
            joint.push_back(wj.second);
        }
        start.push_back((int32_t)joint.size());
    }
    for (int r = 0; r < 3 * V; ++r)
        for (int k = 0; k < K; ++k) key[(size_t)r * K + k] = keyClouds(r, k);
    for (int r = 0; r < 3 * J; ++r)
        for (int k = 0; k < K; ++k) jsr[(size_t)r * K + k] = jointShapeReg(r, k);
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c) faces[3 * (size_t)f + c] = mesh(c, f);
    for (int j = 0; j < J; ++j) par[j] = parent(j);
    avb_model_desc d{};
    d.num_points = V; d.num_joints = J; d.num_shape_keys = K; d.num_faces = F;
    d.base_cloud = baseCloud.data();
    d.key_clouds = key.data();
    d.joint_shape_reg_base = jointShapeRegBase.data();
    d.joint_shape_reg = jsr.data();
    d.parent = par.data();
    d.mesh = faces.data();
    d.assign_start = start.data(); d.assign_joint = joint.data(); d.assign_weight = weight.data();
    if (hasPosePrior()) {
        const int C = posePrior.nComps, D = posePrior.nDims;
        mean.resize((size_t)C * D);
        cov.resize((size_t)C * D * D);
        for (int c = 0; c < C; ++c) {
            for (int j = 0; j < D; ++j) mean[(size_t)c * D + j] = posePrior.mean(c, j);
            for (int j = 0; j < D; ++j)
                for (int k = 0; k < D; ++k) cov[((size_t)c * D + j) * D + k] = posePrior.cov[c](j, k);
        }
        d.gmm_components = C; d.gmm_dims = D;
        d.gmm_weight = posePrior.weight.data(); d.gmm_mean = mean.data(); d.gmm_cov = cov.data();
    }
    if (avb_model_create(&d, &handle_) != AVB_OK) die("avb_model_create");
    return handle_;
}

Avatar::Avatar(const AvatarModel& m) : model(m) {  // Avatar.cpp:12-20
    w.resize(model.numShapeKeys());
    r.resize(model.numJoints());
    w.setZero();
    p.setZero();
    for (auto& R : r) R.setIdentity();
}
Avatar::~Avatar() {
    if (updater_) avb_fitter_destroy(updater_);
}

std::vector<double> Avatar::packParams() const {
    const int J = model.numJoints(), K = model.numShapeKeys();
    std::vector<double> x(3 + 4 * (size_t)J + K);
    for (int c = 0; c < 3; ++c) x[c] = p(c);
    for (int j = 0; j < J; ++j) {
        double R[9];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) R[3 * a + b] = r[j](a, b);
        avb_rotmat_to_quat(R, &x[3 + 4 * j]);
    }
    for (int k = 0; k < K; ++k) x[3 + 4 * J + k] = w(k);
    return x;
}
void Avatar::unpackParams(const std::vector<double>& x) {
    const int J = model.numJoints(), K = model.numShapeKeys();
    for (int c = 0; c < 3; ++c) p(c) = x[c];
    for (int j = 0; j < J; ++j) {
        double R[9];
        avb_quat_to_rotmat(&x[3 + 4 * j], R);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) r[j](a, b) = R[3 * a + b];
    }
    for (int k = 0; k < K; ++k) w(k) = x[3 + 4 * J + k];
}

void Avatar::update() {
    if (!updater_) {
        std::vector<int32_t> pm(model.numJoints(), 0);
        avb_fitter_config cfg{0, 1, 16, 1, pm.data()};
        if (avb_fitter_create(model.handle(), &cfg, &updater_) != AVB_OK) die("avb_fitter_create");
    }
    const std::vector<double> x = packParams();
    cloud.resize(3, model.numPoints());
    jointPos.resize(3, model.numJoints());
    jointTrans.resize(12, model.numJoints());
    if (avb_avatar_update(updater_, 1, x.data(), cloud.data(), jointPos.data(), jointTrans.data()) != AVB_OK)
        die("avb_avatar_update");
}

Eigen::VectorXd Avatar::smplParams() const {
    const int J = model.numJoints();
    Eigen::VectorXd res;
    res.resize((J - 1) * 3);
    const std::vector<double> x = packParams();
    for (int j = 1; j < J; ++j) {
        const double* q = &x[3 + 4 * j];
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        const double s = n == 0 ? 0.0 : 2.0 * std::atan2(n, std::fabs(q[3])) / (q[3] < 0 ? -n : n);
        for (int c = 0; c < 3; ++c) res((j - 1) * 3 + c) = q[c] * s;
    }
    return res;
}

// AvatarOptimizer.cpp:1213-1244
AvatarOptimizer::AvatarOptimizer(Avatar& a, const CameraIntrin& in, const cv::Size& image_size, int num_parts,
                                 const std::vector<int>& part_map)
    : ava(a), intrin(in), imageSize(image_size), numParts(num_parts), partMap(part_map) {
    r.resize(ava.model.numJoints());
}
AvatarOptimizer::~AvatarOptimizer() {
    if (fitter_) avb_fitter_destroy(fitter_);
}

// AvatarOptimizer.cpp:1246-1517
void AvatarOptimizer::optimize(const Eigen::Matrix<double, 3, Eigen::Dynamic>& data_cloud,
                               const Eigen::VectorXi& data_part_labels, int icp_iters, int /*num_threads*/) {
    const int N = (int)data_cloud.cols();
    if (!fitter_ || N > capacity_) {
        if (fitter_) avb_fitter_destroy(fitter_);
        capacity_ = std::max(N, 1 << 16) * 2;
        std::vector<int32_t> pm(partMap.begin(), partMap.end());
        avb_fitter_config cfg{0, 1, capacity_, numParts, pm.data()};
        if (avb_fitter_create(ava.model.handle(), &cfg, &fitter_) != AVB_OK) die("avb_fitter_create");
    }
    std::vector<double> x = ava.packParams();
    avb_options o;
    avb_default_options(&o);
    o.icp_iters = icp_iters;
    o.max_iters_per_icp = maxItersPerICP;
    o.beta_pose = betaPose;
    o.beta_shape = betaShape;
    o.enable_occlusion = enableOcclusion ? 1 : 0;
    o.nn_step = nnStep;
    avb_stats st{};
    ava.cloud.resize(3, ava.model.numPoints());
    std::vector<int32_t> labels(N);
    for (int i = 0; i < N; ++i) labels[i] = data_part_labels(i);
    if (avb_fit(fitter_, data_cloud.data(), labels.data(), N, x.data(), &o, &st, ava.cloud.data()) != AVB_OK)
        die("avb_fit");
    const int J = ava.model.numJoints();
    for (int j = 0; j < J; ++j) r[j] = Eigen::Quaterniond(x[6 + 4 * j], x[3 + 4 * j], x[4 + 4 * j], x[5 + 4 * j]);
    ava.unpackParams(x);                                                  // :1494-1496
    ava.jointPos.resize(3, J);
    ava.jointTrans.resize(12, J);
    if (avb_avatar_update(fitter_, 1, x.data(), nullptr, ava.jointPos.data(), ava.jointTrans.data()) != AVB_OK)
        die("avb_avatar_update");                                         // :1497 (cloud came back from avb_fit)
    lastIterations = st.iterations;
    lastCorrespondences = st.num_correspondences;
    lastInitialCost = st.initial_cost;
    lastFinalCost = st.final_cost;
}
}  // namespace ark

// ---- Avatar::randomize / alignToJoints, GaussianMixture::sample (host-side utilities of the reference's callers) ---------------
namespace ark {
namespace {
// dense lower Cholesky of a small SPD matrix (Eigen::LLT in the reference, GaussianMixture.cpp:44-63)
std::vector<double> chol_lower_dense(const Eigen::MatrixXd& A) {
    const long n = A.rows();
    std::vector<double> L((size_t)n * n, 0.0);
    for (long j = 0; j < n; ++j) {
        double d = A(j, j);
        for (long k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) throw std::runtime_error("Decomposition failed!");
        d = std::sqrt(d);
        L[j * n + j] = d;
        for (long i = j + 1; i < n; ++i) {
            double acc = A(i, j);
            for (long k = 0; k < j; ++k) acc -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = acc / d;
        }
    }
    return L;
}
void axis_angle_to_rot(const double* aa, Eigen::Matrix3d& R) {   // AngleAxisd(angle, axis).toRotationMatrix()
    const double ang = std::sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
    if (ang == 0.0) { R.setIdentity(); return; }
    const double x = aa[0] / ang, y = aa[1] / ang, z = aa[2] / ang, c = std::cos(ang), sn = std::sin(ang), t = 1.0 - c;
    R(0, 0) = t * x * x + c;      R(0, 1) = t * x * y - sn * z; R(0, 2) = t * x * z + sn * y;
    R(1, 0) = t * x * y + sn * z; R(1, 1) = t * y * y + c;      R(1, 2) = t * y * z - sn * x;
    R(2, 0) = t * x * z - sn * y; R(2, 1) = t * y * z + sn * x; R(2, 2) = t * z * z + c;
}
void mat3_mul(const Eigen::Matrix3d& A, const Eigen::Matrix3d& B, Eigen::Matrix3d& C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
}
// Eigen::Quaterniond::FromTwoVectors(a, b).toRotationMatrix(): the shortest rotation taking a to b
void rot_from_two_vectors(const double* a, const double* b, Eigen::Matrix3d& R) {
    const double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    double v0[3] = {a[0] / na, a[1] / na, a[2] / na}, v1[3] = {b[0] / nb, b[1] / nb, b[2] / nb};
    const double c = v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2];
    double q[4];   // x y z w
    if (c < -1.0 + 1e-12) {   // opposite vectors: any axis orthogonal to v0 (Eigen takes it from an SVD; same rotation angle pi)
        double ax[3] = {0, -v0[2], v0[1]};
        if (std::fabs(v0[0]) > 0.9) { ax[0] = -v0[2]; ax[1] = 0; ax[2] = v0[0]; }
        const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        q[0] = ax[0] / n; q[1] = ax[1] / n; q[2] = ax[2] / n; q[3] = 0.0;
    } else {
        const double axis[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
        const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
        q[0] = axis[0] * invs; q[1] = axis[1] * invs; q[2] = axis[2] * invs; q[3] = s * 0.5;
    }
    double Rm[9];
    avb_quat_to_rotmat(q, Rm);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R(i, j) = Rm[3 * i + j];
}
std::mt19937& facade_rng() {
    thread_local static std::mt19937 rg(std::random_device{}());
    return rg;
}
}  // namespace

// GaussianMixture.cpp:116-132
Eigen::VectorXd GaussianMixture::sample() const {
    std::mt19937& rg = facade_rng();
    float randf = std::uniform_real_distribution<float>(0.0f, 1.0f)(rg);
    int component = 0;
    for (int i = 0; i < nComps; ++i) {   // the reference keeps the LAST component whose running weight passes randf
        randf -= (float)weight(i);
        if (randf <= 0) component = i;
    }
    Eigen::VectorXd rv(nDims);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<double> z(nDims);
    for (int i = 0; i < nDims; ++i) z[i] = nd(rg);
    const std::vector<double> L = chol_lower_dense(cov[component]);
    for (int i = 0; i < nDims; ++i) {     // r *= cov_cho is r^T L in the reference (row vector times lower factor)
        double acc = 0;
        for (int k = i; k < nDims; ++k) acc += z[k] * L[(size_t)k * nDims + i];
        rv(i) = acc + mean(component, i);
    }
    return rv;
}

// Avatar.cpp:77-126
void Avatar::randomize(bool randomize_pose, bool randomize_shape, bool randomize_root_pos_rot, uint32_t seed) {
    thread_local static std::mt19937 rg(std::random_device{}());
    if (~seed) rg.seed(seed);
    auto randn = [&](float mean, float sd) { return std::normal_distribution<float>(mean, sd)(rg); };
    auto uniform = [&](float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(rg); };
    if (randomize_shape)
        for (int i = 0; i < model.numShapeKeys(); ++i) w(i) = randn(0.f, 1.f);
    if (randomize_pose && model.hasPosePrior()) {
        const Eigen::VectorXd samp = model.posePrior.sample();
        for (int i = 0; i < model.numJoints() - 1; ++i) {
            const double aa[3] = {samp(3 * i), samp(3 * i + 1), samp(3 * i + 2)};
            axis_angle_to_rot(aa, r[i + 1]);
        }
    }
    if (randomize_root_pos_rot) {
        p(0) = uniform(-1.0f, 1.0f);
        p(1) = uniform(-0.5f, 0.5f);
        p(2) = uniform(2.2f, 4.5f);
        const double angle_up = uniform((float)(-M_PI / 3.), (float)(M_PI / 3.)) + M_PI;
        const double theta = uniform(0.f, (float)(2 * M_PI)), phi = uniform((float)(-M_PI / 2), (float)(M_PI / 2));
        const double axis_perturb[3] = {std::sin(phi) * std::cos(theta), std::cos(phi), std::sin(phi) * std::sin(theta)};   // fromSpherical
        const double angle_perturb = randn(0.0f, 0.2f);
        const double up[3] = {0.0, angle_up, 0.0};
        const double pert[3] = {axis_perturb[0] * angle_perturb, axis_perturb[1] * angle_perturb, axis_perturb[2] * angle_perturb};
        Eigen::Matrix3d Rup, Rp;
        axis_angle_to_rot(up, Rup);
        axis_angle_to_rot(pert, Rp);
        mat3_mul(Rp, Rup, r[0]);   // (aa_perturb * aa_up).toRotationMatrix()
    }
}

// Avatar.cpp:141-193
void Avatar::alignToJoints(const CloudType& pos) {
    const int J = model.numJoints();
    if (pos.cols() != J) throw std::runtime_error("alignToJoints: need one position per joint");
    auto jp = [&](const CloudType& m, int j, double* out) { out[0] = m(0, j); out[1] = m(1, j); out[2] = m(2, j); };
    auto sub = [](const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; };
    auto nrm = [](const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    const int SPINE1 = 3, SPINE2 = 6, ROOT = 0;
    double a0[3], a1[3], b0[3], b1[3], vr[3], vrt[3];
    jp(model.initialJointPos, SPINE1, a1); jp(model.initialJointPos, ROOT, a0); sub(a1, a0, vr);
    jp(pos, SPINE1, b1); jp(pos, ROOT, b0); sub(b1, b0, vrt);
    if (!std::isnan(pos(0, 0))) for (int c = 0; c < 3; ++c) p(c) = pos(c, 0);
    if (!std::isnan(vr[0]) && !std::isnan(vrt[0])) rot_from_two_vectors(vr, vrt, r[0]);
    else r[0].setIdentity();
    std::vector<Eigen::Matrix3d> rotTrans(J);
    rotTrans[0] = r[0];
    double scaleAvg = 0.0;
    for (int i = 1; i < J; ++i) {
        double u[3], v[3], x0[3], x1[3];
        jp(pos, i, x1); jp(pos, model.parent(i), x0); sub(x1, x0, u);
        jp(model.initialJointPos, i, x1); jp(model.initialJointPos, model.parent(i), x0); sub(x1, x0, v);
        scaleAvg += nrm(u) / nrm(v);
    }
    scaleAvg /= (J - 1.0);
    double s2[3], s0[3], sv[3];
    jp(model.initialJointPos, SPINE2, s2); jp(model.initialJointPos, ROOT, s0); sub(s2, s0, sv);
    const double baseScale = nrm(sv) * (scaleAvg - 1.0);
    const double PC1_DIST_FACT = 32.0;
    if (model.numShapeKeys() > 0) {
        w(0) = baseScale * PC1_DIST_FACT;
        if (std::isnan(w(0))) w(0) = 1.5;
    }
    for (int i = 1; i < J; ++i) {
        rotTrans[i] = rotTrans[model.parent(i)];
        if (!std::isnan(pos(0, i))) {
            double vv[3], vvt[3], x0[3], x1[3];
            jp(model.initialJointPos, i, x1); jp(model.initialJointPos, model.parent(i), x0); sub(x1, x0, vv);
            jp(pos, i, x1); jp(pos, model.parent(i), x0); sub(x1, x0, vvt);
            rot_from_two_vectors(vv, vvt, rotTrans[i]);
            const Eigen::Matrix3d& Pm = rotTrans[model.parent(i)];
            for (int a = 0; a < 3; ++a)      // r[i] = rotTrans[parent]^T * rotTrans[i]
                for (int b = 0; b < 3; ++b) r[i](a, b) = Pm(0, a) * rotTrans[i](0, b) + Pm(1, a) * rotTrans[i](1, b) + Pm(2, a) * rotTrans[i](2, b);
        } else {
            r[i].setIdentity();
        }
    }
}

// ---- ark::AvatarRenderer (AvatarRenderer.cpp) over avb_render_batch / avb_render_lambert_batch ------------------------------------
AvatarRenderer::AvatarRenderer(const Avatar& a, const CameraIntrin& in) : ava(a), intrin(in) {}
AvatarRenderer::~AvatarRenderer() {
    if (fitter_) avb_fitter_destroy(fitter_);
}
avb_fitter* AvatarRenderer::fitter(const std::vector<int>& part_map) const {
    std::vector<int> pm = part_map;
    if (pm.empty()) pm.assign(ava.model.numJoints(), 0);
    if (fitter_ && pm == fitter_part_map_) return fitter_;
    if (fitter_) avb_fitter_destroy(fitter_);
    fitter_ = nullptr;
    int nparts = 1;
    for (int v : pm) nparts = std::max(nparts, v + 1);
    std::vector<int32_t> pm32(pm.begin(), pm.end());
    avb_fitter_config cfg{0, 1, 16, nparts, pm32.data()};
    if (avb_fitter_create(ava.model.handle(), &cfg, &fitter_) != AVB_OK) die("avb_fitter_create");
    fitter_part_map_ = pm;
    return fitter_;
}
cv::Mat AvatarRenderer::renderDepth(const cv::Size& sz) const {
    cv::Mat out(sz, CV_32F);
    const std::vector<double> x = ava.packParams();
    avb_render_desc d{sz.width, sz.height, intrin.fx, intrin.cx, intrin.fy, intrin.cy};
    if (avb_render_batch(fitter({}), 1, x.data(), &d, out.ptr<float>(0), nullptr, nullptr) != AVB_OK) die("avb_render_batch");
    return out;
}
cv::Mat AvatarRenderer::renderLambert(const cv::Size& sz) const {
    cv::Mat out(sz, CV_8U);
    const std::vector<double> x = ava.packParams();
    avb_render_desc d{sz.width, sz.height, intrin.fx, intrin.cx, intrin.fy, intrin.cy};
    if (avb_render_lambert_batch(fitter({}), 1, x.data(), &d, out.ptr<uint8_t>(0)) != AVB_OK) die("avb_render_lambert_batch");
    return out;
}
cv::Mat AvatarRenderer::renderPartMask(const cv::Size& sz, const std::vector<int>& part_map) const {
    cv::Mat out(sz, CV_8U);
    const std::vector<double> x = ava.packParams();
    avb_render_desc d{sz.width, sz.height, intrin.fx, intrin.cx, intrin.fy, intrin.cy};
    if (avb_render_batch(fitter(part_map), 1, x.data(), &d, nullptr, out.ptr<uint8_t>(0), nullptr) != AVB_OK) die("avb_render_batch");
    return out;
}
cv::Mat AvatarRenderer::renderFaces(const cv::Size& sz, int) const {
    cv::Mat out(sz, CV_32S);
    const std::vector<double> x = ava.packParams();
    avb_render_desc d{sz.width, sz.height, intrin.fx, intrin.cx, intrin.fy, intrin.cy};
    if (avb_render_batch(fitter({}), 1, x.data(), &d, nullptr, nullptr, out.ptr<int32_t>(0)) != AVB_OK) die("avb_render_batch");
    return out;
}
}  // namespace ark

// ---- ark::RTree data side (RTree.cpp:2967-3061, 3452-3510) ----------------------------------------------------------
namespace ark {
namespace {
template <class T> bool read_bin(std::istream& is, T& v) { return (bool)is.read(reinterpret_cast<char*>(&v), sizeof(T)); }
}

bool RTree::loadFile(const std::string& path) {
    std::ifstream bifs(path, std::ios::in | std::ios::binary);
    if (!bifs) return false;
    char marker = 0;
    bifs.get(marker);
    if (marker == 'R') {   // binary format
        uint32_t nNodes = 0, nLeafs = 0;
        int32_t np = 0;
        if (!read_bin(bifs, nNodes) || !read_bin(bifs, nLeafs) || !read_bin(bifs, np)) throw std::runtime_error("RTree file is truncated: " + path);
        numParts = np;
        nodes.assign(nNodes, RNode());
        leafData.assign(nLeafs, std::vector<float>());
        uint32_t last = 0;
        for (size_t i = 0; i < nodes.size(); ++i) {
            uint8_t isLeaf = 0;
            if (!read_bin(bifs, isLeaf)) throw std::runtime_error("RTree file is truncated: " + path);
            if (isLeaf) {
                if (last >= nLeafs) throw std::runtime_error("RTree file has more leaves than announced: " + path);
                leafData[last].assign(numParts, 0.f);
                uint8_t cnt = 0;
                read_bin(bifs, cnt);
                if (cnt > numParts) throw std::runtime_error("RTree leaf has more parts than numParts: " + path);
                for (uint8_t j = 0; j < cnt; ++j) {
                    uint8_t k = 0;
                    float val = 0;
                    read_bin(bifs, k);
                    if (k >= numParts || !read_bin(bifs, val)) throw std::runtime_error("RTree leaf entry out of bounds: " + path);
                    leafData[last][k] = val;
                }
                nodes[i].leafid = (int)last++;
            } else {
                int32_t l = 0, r = 0;
                read_bin(bifs, l);
                read_bin(bifs, r);
                nodes[i].lnode = l;
                nodes[i].rnode = r;
                read_bin(bifs, nodes[i].thresh);
                bifs.read(reinterpret_cast<char*>(nodes[i].u), sizeof(float) * 2);
                if (!bifs.read(reinterpret_cast<char*>(nodes[i].v), sizeof(float) * 2)) throw std::runtime_error("RTree file is truncated: " + path);
            }
        }
        bifs.get(marker);
        if (marker != 'T') throw std::runtime_error("incorrect RTree format, T end marker missing: " + path);
    } else {               // legacy text format
        bifs.close();
        std::ifstream ifs(path);
        if (!ifs) return false;
        size_t nNodes = 0, nLeafs = 0;
        ifs >> nNodes >> nLeafs >> numParts;
        if (!ifs) throw std::runtime_error("RTree text file has no header: " + path);
        nodes.assign(nNodes, RNode());
        leafData.assign(nLeafs, std::vector<float>());
        for (size_t i = 0; i < nNodes; ++i) {
            ifs >> nodes[i].leafid;
            if (nodes[i].leafid < 0)
                ifs >> nodes[i].lnode >> nodes[i].rnode >> nodes[i].thresh >> nodes[i].u[0] >> nodes[i].u[1] >> nodes[i].v[0] >> nodes[i].v[1];
        }
        for (size_t i = 0; i < nLeafs; ++i) {
            leafData[i].assign(numParts, 0.f);
            for (int j = 0; j < numParts; ++j) ifs >> leafData[i][j];
        }
        if (!ifs) throw std::runtime_error("RTree text file is truncated: " + path);
    }
    updateBestMatchTable();
    int numNew = 0;
    if (!readPartMap(path + ".partmap", partMap, numNew, partMapType))
        std::fprintf(stderr, "Warning: partmap not found or invalid beside %s; using default map\n", path.c_str());
    return true;
}

void RTree::updateBestMatchTable() {
    leafBestMatch.assign(leafData.size(), 0);
    for (size_t i = 0; i < leafData.size(); ++i) {
        float best = std::numeric_limits<float>::lowest();
        for (int j = 0; j < numParts && j < (int)leafData[i].size(); ++j)
            if (leafData[i][j] > best) {
                best = leafData[i][j];
                leafBestMatch[i] = (uint8_t)j;
            }
    }
}

bool RTree::readPartMap(const std::string& path, std::vector<int>& result, int& num_new_parts, int& partmap_type) {
    std::ifstream is(path);
    if (!is) return false;
    std::string marker;
    is >> marker;
    if (marker != "partmap") return false;
    is >> marker;
    if (marker == "disjoint") partmap_type = 1;
    else if (marker == "contiguous") partmap_type = 0;
    else return false;
    int nOld = 0, nNew = 0;
    is >> marker;
    if (marker != "src") return false;
    is >> nOld;
    std::map<std::string, int> oldEnum, newEnum;
    for (int i = 0; i < nOld; ++i) { std::string name; is >> name; oldEnum[name] = i; }
    is >> marker;
    if (marker != "dest") return false;
    is >> nNew;
    for (int i = 0; i < nNew; ++i) { std::string name; is >> name; newEnum[name] = i; }
    result.assign(nOld, 0);
    for (int i = 0; i < nOld; ++i) {
        std::string a, b;
        if (!(is >> a >> b)) break;
        result[oldEnum[a]] = newEnum[b];
    }
    num_new_parts = nNew;
    return true;
}

void RTree::attach(avb_fitter* fitter) const {
    std::vector<float> u(2 * nodes.size()), v(2 * nodes.size()), th(nodes.size());
    std::vector<int32_t> l(nodes.size()), r(nodes.size()), id(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        u[2 * i] = nodes[i].u[0]; u[2 * i + 1] = nodes[i].u[1]; v[2 * i] = nodes[i].v[0]; v[2 * i + 1] = nodes[i].v[1];
        th[i] = nodes[i].thresh; l[i] = nodes[i].lnode; r[i] = nodes[i].rnode; id[i] = nodes[i].leafid;
    }
    avb_rtree_desc d{(int32_t)nodes.size(), (int32_t)leafBestMatch.size(), numParts, u.data(), v.data(), th.data(), l.data(), r.data(),
                     id.data(), leafBestMatch.data()};
    if (avb_fitter_set_rtree(fitter, &d) != AVB_OK) die("avb_fitter_set_rtree");
}

RTree::~RTree() {
    if (fitter_) avb_fitter_destroy(fitter_);
    if (dummy_) avb_model_destroy(dummy_);
}
// the RTree kernels need no body model; the C ABI hangs them on a fitter, so the tree owns one on a one-triangle model
avb_fitter* RTree::device(size_t pixels) const {
    if (!dummy_) {
        static const double base[9] = {0, 0, 1, 1, 0, 1, 0, 1, 1}, jbase[3] = {0, 0, 1}, wgt[3] = {1, 1, 1};
        static const int32_t parent[1] = {-1}, mesh[3] = {0, 1, 2}, start[4] = {0, 1, 2, 3}, joint[3] = {0, 0, 0};
        avb_model_desc d{};
        d.num_points = 3; d.num_joints = 1; d.num_shape_keys = 0; d.num_faces = 1;
        d.base_cloud = base; d.joint_shape_reg_base = jbase; d.parent = parent; d.mesh = mesh;
        d.assign_start = start; d.assign_joint = joint; d.assign_weight = wgt;
        if (avb_model_create(&d, &dummy_) != AVB_OK) die("avb_model_create");
    }
    if (!fitter_) {
        const int32_t pm[1] = {0};
        avb_fitter_config cfg{0, 1, (int64_t)std::max<size_t>(pixels, 1 << 20), std::max(numParts, 1), pm};
        cfg.num_parts = 1;
        if (avb_fitter_create(dummy_, &cfg, &fitter_) != AVB_OK) die("avb_fitter_create");
        attached_nodes_ = 0;
    }
    if (attached_nodes_ != nodes.size() && !nodes.empty()) {
        attach(fitter_);
        attached_nodes_ = nodes.size();
    }
    return fitter_;
}
cv::Mat RTree::predictBest(const cv::Mat& depth, int, int interval, cv::Point top_left, cv::Point bot_right, bool fill_in_gaps) {
    cv::Mat result(depth.size(), CV_8U);
    const bool whole = bot_right.x == -1;
    const int32_t roi[4] = {top_left.x, top_left.y, whole ? depth.cols - 1 : bot_right.x, whole ? depth.rows - 1 : bot_right.y};
    if (avb_rtree_predict_batch(device((size_t)depth.rows * depth.cols), 1, depth.ptr<float>(0), depth.cols, depth.rows, roi, interval,
                                fill_in_gaps ? 1 : 0, result.ptr<uint8_t>(0)) != AVB_OK)
        die("avb_rtree_predict_batch");
    return result;
}
void RTree::postProcess(cv::Mat& image, Eigen::Matrix<double, 2, Eigen::Dynamic>& com_pre, int interval, int, cv::Point top_left,
                        cv::Point bot_right, double dist_to_pre_weight) const {
    if (bot_right.x == -1) {
        bot_right.x = image.cols - 1;
        bot_right.y = image.rows - 1;
    }
    if (com_pre.cols() != numParts) {   // RTree.cpp:3432-3436
        com_pre.resize(2, numParts);
        for (int i = 0; i < numParts; ++i) { com_pre(0, i) = -1.; com_pre(1, i) = 0.; }
    }
    const int32_t roi[4] = {top_left.x, top_left.y, bot_right.x, bot_right.y};
    if (avb_rtree_postprocess_batch(device((size_t)image.rows * image.cols), 1, image.ptr<uint8_t>(0), image.cols, image.rows, roi, interval,
                                    numParts, partMapType == 0 ? 0 : 1, com_pre.data(), dist_to_pre_weight) != AVB_OK)
        die("avb_rtree_postprocess_batch");
}

}  // namespace ark

// test hook (tests/test_cpp_facade.py): load a tree file with the facade and copy its arrays out
extern "C" int ark_b200_rtree_probe(const char* path, int32_t* counts3, float* uvt /*[nodes][5]*/, int32_t* lri /*[nodes][3]*/,
                                    uint8_t* leaf_best, int32_t cap_nodes, int32_t cap_leaves, int32_t* partmap_info /*[2 + 64]*/) {
    try {
        ark::RTree t(0);
        if (!t.loadFile(path)) return 1;
        counts3[0] = (int32_t)t.nodes.size(); counts3[1] = (int32_t)t.leafBestMatch.size(); counts3[2] = t.numParts;
        if ((int)t.nodes.size() > cap_nodes || (int)t.leafBestMatch.size() > cap_leaves) return 2;
        for (size_t i = 0; i < t.nodes.size(); ++i) {
            uvt[5 * i] = t.nodes[i].u[0]; uvt[5 * i + 1] = t.nodes[i].u[1]; uvt[5 * i + 2] = t.nodes[i].v[0]; uvt[5 * i + 3] = t.nodes[i].v[1];
            uvt[5 * i + 4] = t.nodes[i].thresh;
            lri[3 * i] = t.nodes[i].lnode; lri[3 * i + 1] = t.nodes[i].rnode; lri[3 * i + 2] = t.nodes[i].leafid;
        }
        for (size_t i = 0; i < t.leafBestMatch.size(); ++i) leaf_best[i] = t.leafBestMatch[i];
        partmap_info[0] = t.partMapType;
        partmap_info[1] = (int32_t)t.partMap.size();
        for (size_t i = 0; i < t.partMap.size() && i < 64; ++i) partmap_info[2 + i] = t.partMap[i];
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ark_b200_rtree_probe: %s\n", e.what());
        return 3;
    }
}

// ---- ark::CameraIntrin (Calibration.cpp:13-51, 68-80, 97-111); only without OpenCV, like its declaration --------------
#if !__has_include(<opencv2/core.hpp>)
namespace ark {
void CameraIntrin::clear() {
    fx = fy = cx = cy = 0.f;
    std::fill(k, k + 6, 0.f);
    std::fill(p, p + 2, 0.f);
}
bool CameraIntrin::readFile(const std::string& path) {
    clear();
    std::ifstream ifs(path);
    int good = 0;
    while (ifs) {
        std::string tag;
        ifs >> tag;
        if (tag.size() != 2) continue;
        if (tag == "cx") { ifs >> cx; ++good; }
        else if (tag == "cy") { ifs >> cy; ++good; }
        else if (tag == "fx") { ifs >> fx; ++good; }
        else if (tag == "fy") { ifs >> fy; ++good; }
        else if (tag[0] == 'k') {
            const int idx = tag[1] - '1';
            if (idx < 0 || idx >= 6) continue;
            ifs >> k[idx];
        } else if (tag[0] == 'p') {
            const int idx = tag[1] - '1';
            if (idx < 0 || idx >= 6) continue;   // the reference checks against 6 here too (p has 2 entries)
            if (idx < 2) ifs >> p[idx];
            else { float skip; ifs >> skip; }
        }
    }
    return good == 4;
}
bool CameraIntrin::writeFile(const std::string& path) const {
    std::ofstream ofs(path);
    if (!ofs) return false;
    ofs << "fx " << fx << "\ncx " << cx << "\nfy " << fy << "\ncy " << cy << "\n";
    for (int i = 0; i < 6; ++i)
        if (k[i] != 0.f) ofs << "k" << i << " " << k[i] << "\n";   // 0-based tags, as the reference writes them
    for (int i = 0; i < 2; ++i)
        if (p[i] != 0.f) ofs << "p" << i << " " << p[i] << "\n";
    return true;
}
void CameraIntrin::to3D(float px, float py, float depth, float out[3]) const {
    out[0] = (px - cx) * depth / fx;
    out[1] = (py - cy) * depth / fy;
    out[2] = depth;
}
void CameraIntrin::to2D(const float xyz[3], float out[2]) const {
    out[0] = xyz[0] * fx / xyz[2] + cx;
    out[1] = xyz[1] * fy / xyz[2] + cy;
}
}  // namespace ark

// test hook: read an intrinsics file with the facade, write it back, return the twelve numbers
extern "C" int ark_b200_intrin_probe(const char* path_in, const char* path_out, float* out12) {
    ark::CameraIntrin c;
    const bool ok = c.readFile(path_in);
    out12[0] = c.fx; out12[1] = c.fy; out12[2] = c.cx; out12[3] = c.cy;
    for (int i = 0; i < 6; ++i) out12[4 + i] = c.k[i];
    out12[10] = c.p[0]; out12[11] = c.p[1];
    if (path_out && !c.writeFile(path_out)) return 2;
    return ok ? 0 : 1;
}
#endif
