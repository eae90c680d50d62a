// oracle/shim/ceres/ceres.h -- TEST INFRASTRUCTURE, not Ceres.  The declarations of Ceres Solver's public interface that the
// reference's AvatarOptimizer.cpp names (CostFunction, LocalParameterization, EvaluationCallback, Problem, Solver::Options /
// Summary, Solve), so that the reference source compiles unchanged into oracle/_ref.  Problem only RECORDS the blocks;
// ceres::Solve is defined by oracle/ref_optimizer.cpp, which evaluates the reference's own cost functors and drives them
// with a restated Levenberg-Marquardt loop (Ceres itself is not in this image: the solver policy is NOT pinned by this).
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#define CERES_VERSION_MAJOR 2

namespace ceres {

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    const std::vector<int>& parameter_block_sizes() const { return sizes_; }
    int num_residuals() const { return nres_; }
protected:
    std::vector<int>* mutable_parameter_block_sizes() { return &sizes_; }
    void set_num_residuals(int n) { nres_ = n; }
private:
    std::vector<int> sizes_;
    int nres_ = 0;
};

class LossFunction;

class LocalParameterization {
public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};

class EvaluationCallback {
public:
    virtual ~EvaluationCallback() {}
    virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum PreconditionerType { IDENTITY, JACOBI, SCHUR_JACOBI, CLUSTER_JACOBI, CLUSTER_TRIDIAGONAL };
enum DoglegType { TRADITIONAL_DOGLEG, SUBSPACE_DOGLEG };
enum LoggingType { SILENT, PER_MINIMIZER_ITERATION };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum LineSearchDirectionType { STEEPEST_DESCENT, NONLINEAR_CONJUGATE_GRADIENT, LBFGS, BFGS };
enum LineSearchInterpolationType { BISECTION, QUADRATIC, CUBIC };
enum LineSearchType { ARMIJO, WOLFE };
enum DenseLinearAlgebraLibraryType { EIGEN, LAPACK };
typedef void* ResidualBlockId;

class Problem {
public:
    struct Options {
        EvaluationCallback* evaluation_callback = nullptr;
    };
    struct EvaluateOptions {
        std::vector<ResidualBlockId> residual_blocks;
    };
    struct ParamBlock {
        double* ptr;
        int size;
        LocalParameterization* local;
        bool constant;
    };
    struct ResidualBlock {
        CostFunction* cost;
        std::vector<double*> params;
    };
    Problem() {}
    explicit Problem(const Options& o) : options(o) {}
    ~Problem() {
        for (auto& r : residuals) delete r.cost;
        std::vector<LocalParameterization*> seen;
        for (auto& p : params) {
            bool dup = false;
            for (auto* s : seen) dup = dup || s == p.local;
            if (p.local && !dup) { seen.push_back(p.local); delete p.local; }
        }
    }
    void AddParameterBlock(double* values, int size, LocalParameterization* lp = nullptr) { params.push_back({values, size, lp, false}); }
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction*, const std::vector<double*>& blocks) {
        residuals.push_back({cost, blocks});
        return cost;
    }
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction*, double* x0) {
        residuals.push_back({cost, std::vector<double*>(1, x0)});
        return cost;
    }
    void SetParameterBlockConstant(double* values) {
        for (auto& p : params) if (p.ptr == values) p.constant = true;
    }
    Options options;
    std::vector<ParamBlock> params;
    std::vector<ResidualBlock> residuals;
};

class Solver {
public:
    struct Options {
        LinearSolverType linear_solver_type = DENSE_NORMAL_CHOLESKY;
        TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
        PreconditionerType preconditioner_type = JACOBI;
        DoglegType dogleg_type = TRADITIONAL_DOGLEG;
        double initial_trust_region_radius = 1e4;
        bool minimizer_progress_to_stdout = false;
        LoggingType logging_type = SILENT;
        MinimizerType minimizer_type = TRUST_REGION;
        bool use_approximate_eigenvalue_bfgs_scaling = false;
        bool check_gradients = false;
        LineSearchDirectionType line_search_direction_type = LBFGS;
        LineSearchInterpolationType line_search_interpolation_type = CUBIC;
        LineSearchType line_search_type = WOLFE;
        int max_num_line_search_direction_restarts = 5;
        int max_num_line_search_step_size_iterations = 20;
        int max_linear_solver_iterations = 500;
        int max_num_iterations = 50;
        int num_threads = 1;
        double function_tolerance = 1e-6;
        double gradient_tolerance = 1e-10;
        double parameter_tolerance = 1e-8;
        DenseLinearAlgebraLibraryType dense_linear_algebra_library_type = EIGEN;
        EvaluationCallback* evaluation_callback = nullptr;
    };
    struct Summary {
        int num_successful_steps = 0, num_unsuccessful_steps = 0, iterations = 0;
        double initial_cost = 0, final_cost = 0;
        std::string FullReport() const { return "ceres shim"; }
        std::string BriefReport() const { return "ceres shim"; }
    };
};

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);

}  // namespace ceres
