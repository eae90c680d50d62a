// oracle/ref_optimizer.cpp -- TEST INFRASTRUCTURE.  Runs the REFERENCE's own AvatarOptimizer.cpp (compiled from
// /root/reference by oracle/Makefile against the Eigen / Ceres / boost stand-ins under oracle/shim, never copied): its
// prologue (:1247-1311), back-face visibility (:1347-1362), findNN over its vendored nanoflann (:841-920), its cost functors
// (AvatarICPCostFunctor / AvatarPosePriorCostFunctor / AvatarShapePriorCostFunctor, :463-723), its evaluation callback
// (AvatarEvaluationCommonData, :162-460) and its FakeQuaternionParameterization (:120-154).
//
// What is NOT the reference here is the solver: Ceres is absent, so ceres::Solve below is this file's.  It can
//   (a) CAPTURE: evaluate the reference's residual blocks at the current point and hand back cost, gradient and J^T J in the
//       tangent space of the reference's own local parameterization -- this is what pins the oracle's evaluate(); or
//   (b) run a Levenberg-Marquardt loop with Ceres 1.14's trust-region policy (the same restatement as the oracle's
//       solve_gn_lm) over the reference's residual blocks -- so that a whole AvatarOptimizer::optimize() of reference code
//       can be compared with the oracle's gn_lm fit and with the device.  The reference itself configures Ceres'
//       LINE_SEARCH / BFGS minimizer (:1322-1326); that policy is not reproduced here and stays unpinned.
#include <ceres/ceres.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "Avatar.h"
#include "AvatarOptimizer.h"
#include "AvatarRenderer.h"
#include "Calibration.h"

namespace {
struct Capture {
    int mode = 0;             // 0: LM, 1: capture at the start point and return
    int max_iters = -1;       // > = 0 overrides options.max_num_iterations
    double function_tolerance = -1;
    double cost = 0;
    std::vector<double> grad, H;
    int iterations = 0, accepted = 0, num_residual_blocks = 0;
    double initial_cost = 0, final_cost = 0;
} g_cap;

struct Layout {
    std::vector<int> goff, toff, gsize, tsize;   // per parameter block: offsets / sizes in global and tangent space
    int nx = 0, P = 0;
};

bool cholesky_lower(const double* A, int n, double* L) {
    std::fill(L, L + (size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double s = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) s -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        const double ljj = std::sqrt(s);
        L[(size_t)j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double t = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) t -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
            L[(size_t)i * n + j] = t / ljj;
        }
    }
    return true;
}
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
}  // namespace

namespace ceres {

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
    Layout lay;
    const int nb = (int)problem->params.size();
    for (int b = 0; b < nb; ++b) {
        const auto& pb = problem->params[(size_t)b];
        lay.goff.push_back(lay.nx);
        lay.toff.push_back(lay.P);
        lay.gsize.push_back(pb.size);
        lay.tsize.push_back(pb.local ? pb.local->LocalSize() : pb.size);
        lay.nx += pb.size;
        lay.P += lay.tsize.back();
    }
    const int P = lay.P, nx = lay.nx;
    auto block_of = [&](const double* ptr) {
        for (int b = 0; b < nb; ++b)
            if (problem->params[(size_t)b].ptr == ptr) return b;
        return -1;
    };
    EvaluationCallback* cb = problem->options.evaluation_callback ? problem->options.evaluation_callback : options.evaluation_callback;

    // cost, gradient and J^T J (tangent space) at the CURRENT values of the parameter blocks
    auto evaluate = [&](double* grad, double* H) -> double {
        if (cb) cb->PrepareForEvaluation(true, true);
        std::fill(grad, grad + P, 0.0);
        std::fill(H, H + (size_t)P * P, 0.0);
        double cost = 0.0;
        std::vector<double> res, jl;
        std::vector<std::vector<double>> jac;
        std::vector<double*> jptr;
        std::vector<int> blocks;
        for (const auto& rb : problem->residuals) {
            const int nr = rb.cost->num_residuals();
            const auto& sizes = rb.cost->parameter_block_sizes();
            const int npb = (int)rb.params.size();
            res.assign((size_t)nr, 0.0);
            jac.resize((size_t)npb);
            jptr.resize((size_t)npb);
            blocks.resize((size_t)npb);
            for (int i = 0; i < npb; ++i) {
                jac[(size_t)i].assign((size_t)nr * sizes[(size_t)i], 0.0);
                jptr[(size_t)i] = jac[(size_t)i].data();
                blocks[(size_t)i] = block_of(rb.params[(size_t)i]);
            }
            rb.cost->Evaluate(rb.params.data(), res.data(), jptr.data());
            for (int r = 0; r < nr; ++r) cost += 0.5 * res[(size_t)r] * res[(size_t)r];
            // local Jacobian of the residual block: [nr x sum of tangent sizes], columns placed at the tangent offsets
            int width = 0;
            std::vector<int> col0((size_t)npb);
            for (int i = 0; i < npb; ++i) { col0[(size_t)i] = width; width += lay.tsize[(size_t)blocks[(size_t)i]]; }
            jl.assign((size_t)nr * width, 0.0);
            for (int i = 0; i < npb; ++i) {
                const int b = blocks[(size_t)i], gs = lay.gsize[(size_t)b], ts = lay.tsize[(size_t)b];
                const auto& pb = problem->params[(size_t)b];
                if (pb.local) {
                    std::vector<double> Pj((size_t)gs * ts);
                    pb.local->ComputeJacobian(pb.ptr, Pj.data());   // [gs x ts] row major
                    for (int r = 0; r < nr; ++r)
                        for (int c = 0; c < ts; ++c) {
                            double s = 0;
                            for (int k = 0; k < gs; ++k) s += jac[(size_t)i][(size_t)r * gs + k] * Pj[(size_t)k * ts + c];
                            jl[(size_t)r * width + col0[(size_t)i] + c] = s;
                        }
                } else {
                    for (int r = 0; r < nr; ++r)
                        for (int c = 0; c < ts; ++c) jl[(size_t)r * width + col0[(size_t)i] + c] = jac[(size_t)i][(size_t)r * gs + c];
                }
            }
            // the same parameter block may appear more than once in a residual block's list: accumulate by tangent column
            std::vector<int> tcol((size_t)width);
            for (int i = 0; i < npb; ++i)
                for (int c = 0; c < lay.tsize[(size_t)blocks[(size_t)i]]; ++c)
                    tcol[(size_t)(col0[(size_t)i] + c)] = lay.toff[(size_t)blocks[(size_t)i]] + c;
            for (int a = 0; a < width; ++a) {
                double ga = 0;
                for (int r = 0; r < nr; ++r) ga += jl[(size_t)r * width + a] * res[(size_t)r];
                grad[tcol[(size_t)a]] += ga;
                for (int bcol = 0; bcol < width; ++bcol) {
                    double h = 0;
                    for (int r = 0; r < nr; ++r) h += jl[(size_t)r * width + a] * jl[(size_t)r * width + bcol];
                    H[(size_t)tcol[(size_t)a] * P + tcol[(size_t)bcol]] += h;
                }
            }
        }
        return cost;
    };
    auto read_x = [&](double* x) {
        for (int b = 0; b < nb; ++b) std::memcpy(x + lay.goff[(size_t)b], problem->params[(size_t)b].ptr, sizeof(double) * lay.gsize[(size_t)b]);
    };
    auto write_x = [&](const double* x) {
        for (int b = 0; b < nb; ++b) std::memcpy(problem->params[(size_t)b].ptr, x + lay.goff[(size_t)b], sizeof(double) * lay.gsize[(size_t)b]);
    };
    auto plus = [&](const double* x, const double* d, double* xp) {
        for (int b = 0; b < nb; ++b) {
            const auto& pb = problem->params[(size_t)b];
            if (pb.local) {
                pb.local->Plus(x + lay.goff[(size_t)b], d + lay.toff[(size_t)b], xp + lay.goff[(size_t)b]);
            } else {
                for (int k = 0; k < pb.size; ++k) xp[lay.goff[(size_t)b] + k] = x[lay.goff[(size_t)b] + k] + d[lay.toff[(size_t)b] + k];
            }
        }
    };

    std::vector<double> g(P), H((size_t)P * P), gt(P), Ht((size_t)P * P), A((size_t)P * P), L((size_t)P * P), delta(P), Hd(P),
        x(nx), xt(nx);
    g_cap.num_residual_blocks = (int)problem->residuals.size();
    double cost = evaluate(g.data(), H.data());
    if (g_cap.mode == 1) {
        g_cap.cost = cost;
        g_cap.grad = g;
        g_cap.H = H;
        if (summary) summary->initial_cost = summary->final_cost = cost;
        return;
    }
    // Levenberg-Marquardt, Ceres 1.14 trust-region policy (same restatement as oracle/avatar_oracle.cpp solve_gn_lm)
    const int max_iters = g_cap.max_iters >= 0 ? g_cap.max_iters : options.max_num_iterations;
    const double function_tolerance = g_cap.function_tolerance >= 0 ? g_cap.function_tolerance : options.function_tolerance;
    read_x(x.data());
    int iterations = 0, accepted_n = 0;
    const double initial_cost = cost;
    double radius = 1e4, decrease_factor = 2.0;
    const double kMinDiag = 1e-6, kMaxDiag = 1e32, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    auto gmax = [&](const std::vector<double>& v) {
        double m = 0;
        for (double e : v) m = std::max(m, std::fabs(e));
        return m;
    };
    bool done = gmax(g) <= gradient_tolerance;
    for (int it = 0; it < max_iters && !done; ++it) {
        ++iterations;
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
            plus(x.data(), delta.data(), xt.data());
            write_x(xt.data());
            const double cost_t = evaluate(gt.data(), Ht.data());
            const double rho = (cost - cost_t) / model_change;
            if (std::isfinite(cost_t) && rho > 1e-3) {
                accepted = true;
                ++accepted_n;
                const double cost_change = cost - cost_t;
                double dn = 0, xn = 0;
                for (int i = 0; i < nx; ++i) {
                    dn += (xt[i] - x[i]) * (xt[i] - x[i]);
                    xn += x[i] * x[i];
                }
                x = xt;
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
            write_x(x.data());
            radius /= decrease_factor;
            decrease_factor *= 2.0;
            if (radius < 1e-32) done = true;
        }
    }
    write_x(x.data());
    g_cap.iterations += iterations;
    g_cap.accepted += accepted_n;
    if (g_cap.initial_cost == 0) g_cap.initial_cost = initial_cost;
    g_cap.final_cost = cost;
    if (summary) {
        summary->iterations = iterations;
        summary->num_successful_steps = accepted_n;
        summary->initial_cost = initial_cost;
        summary->final_cost = cost;
    }
}

}  // namespace ceres

namespace {
struct RefModel2 {   // same layout as RefModel of ref_avatar.cpp: the model loaded by the reference's own AvatarModel.cpp
    ark::AvatarModel m;
    explicit RefModel2(const char* dir) : m(dir) {}
};
struct RefOpt {
    RefModel2* model;
    ark::Avatar ava;
    ark::CameraIntrin intrin;
    std::vector<int> part_map;
    ark::AvatarOptimizer* opt = nullptr;
    explicit RefOpt(RefModel2* rm) : model(rm), ava(rm->m) {}
};
void quiet_update(ark::Avatar& ava) {
    FILE* keep = stdout;
    stdout = fopen("/dev/null", "w");
    ava.update();
    fclose(stdout);
    stdout = keep;
}
}  // namespace

extern "C" {

void* ref_opt_create(void* model, int num_parts, const int32_t* part_map) {
    auto* rm = static_cast<RefModel2*>(model);
    auto* ro = new RefOpt(rm);
    ro->part_map.assign(part_map, part_map + rm->m.numJoints());
    quiet_update(ro->ava);
    ro->opt = new ark::AvatarOptimizer(ro->ava, ro->intrin, cv::Size(640, 576), num_parts, ro->part_map);
    return ro;
}
void ref_opt_free(void* h) {
    auto* ro = static_cast<RefOpt*>(h);
    delete ro->opt;
    delete ro;
}

// x: [p(3) | quaternions x,y,z,w (J) | w(K)], in/out.  mode 0: AvatarOptimizer::optimize with the LM loop above; mode 1:
// capture cost / gradient [P] / J^T J [P x P] of the reference's residual blocks at x (correspondences from the reference's
// own visibility + findNN) and leave x unchanged.  stats: iterations, accepted, residual blocks; costs: initial, final.
int ref_opt_run(void* h, double* x, const double* data, const int32_t* labels, int n, int icp_iters, int max_iters_per_icp,
                double function_tolerance, double beta_pose, double beta_shape, int enable_occlusion, int num_threads, int mode,
                double* cost, double* grad, double* H, int32_t* stats, double* costs) {
    auto* ro = static_cast<RefOpt*>(h);
    ark::Avatar& ava = ro->ava;
    const int J = ava.model.numJoints(), K = ava.model.numShapeKeys();
    for (int i = 0; i < 3; ++i) ava.p[i] = x[i];
    for (int j = 0; j < J; ++j) {
        Eigen::Quaterniond q(x[3 + 4 * j + 3], x[3 + 4 * j], x[3 + 4 * j + 1], x[3 + 4 * j + 2]);
        ava.r[(size_t)j] = q.toRotationMatrix();
    }
    for (int k = 0; k < K; ++k) ava.w[k] = x[3 + 4 * J + k];
    FILE* keep = stdout;
    stdout = fopen("/dev/null", "w");
    ava.update();
    ark::AvatarOptimizer& opt = *ro->opt;
    opt.betaPose = beta_pose;
    opt.betaShape = beta_shape;
    opt.maxItersPerICP = max_iters_per_icp;
    opt.enableOcclusion = enable_occlusion != 0;
    ark::CloudType cloud(3, n);
    Eigen::VectorXi lab(n);
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) cloud(c, i) = data[(size_t)i * 3 + c];
        lab[i] = labels[i];
    }
    g_cap = Capture();
    g_cap.mode = mode;
    g_cap.max_iters = max_iters_per_icp;
    g_cap.function_tolerance = function_tolerance;
    opt.optimize(cloud, lab, mode == 1 ? 1 : icp_iters, num_threads);
    fclose(stdout);
    stdout = keep;
    if (mode == 1) {
        const int P = (int)g_cap.grad.size();
        *cost = g_cap.cost;
        for (int i = 0; i < P; ++i) grad[i] = g_cap.grad[(size_t)i];
        for (size_t i = 0; i < (size_t)P * P; ++i) H[i] = g_cap.H[i];
    } else {
        for (int i = 0; i < 3; ++i) x[i] = ava.p[i];
        for (int j = 0; j < J; ++j)
            for (int c = 0; c < 4; ++c) x[3 + 4 * j + c] = opt.r[(size_t)j].coeffs()(c);
        for (int k = 0; k < K; ++k) x[3 + 4 * J + k] = ava.w[k];
    }
    stats[0] = g_cap.iterations;
    stats[1] = g_cap.accepted;
    stats[2] = g_cap.num_residual_blocks;
    costs[0] = g_cap.initial_cost;
    costs[1] = g_cap.final_cost;
    return 0;
}

// The reference's own AvatarRenderer (AvatarRenderer.cpp: projection, painter's ordering, the four render* functions over the
// painters of AvatarHelpers.cpp) on a posed cloud [V][3]; intrin = fx, cx, fy, cy.  Any output may be null.
void ref_render(void* model, const double* cloud, const float* intrin, int width, int height, const int32_t* part_map,
                float* depth, uint8_t* parts, int32_t* faces, uint8_t* lambert) {
    auto* rm = static_cast<RefModel2*>(model);
    ark::Avatar ava(rm->m);
    const int V = rm->m.numPoints(), J = rm->m.numJoints();
    ava.cloud.resize(3, V);
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < 3; ++c) ava.cloud(c, v) = cloud[(size_t)v * 3 + c];
    ava.jointPos.resize(3, J);
    ark::CameraIntrin K;
    K.clear();
    K.fx = intrin[0]; K.cx = intrin[1]; K.fy = intrin[2]; K.cy = intrin[3];
    ark::AvatarRenderer rend(ava, K);
    const cv::Size sz(width, height);
    const size_t npx = (size_t)width * height;
    if (depth) {
        cv::Mat m = rend.renderDepth(sz);
        std::memcpy(depth, m.data, npx * 4);
    }
    if (parts) {
        std::vector<int> pm(part_map, part_map + J);
        cv::Mat m = rend.renderPartMask(sz, pm);
        std::memcpy(parts, m.data, npx);
    }
    if (faces) {
        cv::Mat m = rend.renderFaces(sz, 1);
        std::memcpy(faces, m.data, npx * 4);
    }
    if (lambert) {
        cv::Mat m = rend.renderLambert(sz);
        std::memcpy(lambert, m.data, npx);
    }
}

}  // extern "C"
