/* oracle/avatar_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C++17, fp64, no Eigen/Ceres) of the sxyu/avatar SMPL-to-point-cloud
 * fitting path (ark::AvatarOptimizer::optimize and what it calls).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (avatar_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: the reference has no tests, golden vectors or fixtures for this path
 *  and cannot be compiled here (Eigen, Ceres, OpenCV, Boost absent), so:
 *   - everything upstream of the solver (forward model, visibility, NN, residuals, Jacobians,
 *     priors) is specified by the reference source and restated line by line; the NN step is
 *     additionally pinned against the reference's own vendored nanoflann.hpp (oracle/_ref);
 *   - the widened rows: the renderer's painters and CameraIntrin::depthToXYZ are pinned against the reference's own
 *     AvatarHelpers.cpp / Calibration.cpp compiled into oracle/_ref with container-only stand-ins (oracle/shim);
 *     RTree::predictBest is restated from source (RTree.cpp needs the full Eigen/OpenCV stack);
 *   - the solver trajectory (Ceres 1.14, external, un-vendored) is PARITY UNPINNED; everything the solver calls
 *     (residuals, Jacobians, priors, parameterization, visibility, NN, Avatar::update) is pinned against the
 *     reference's own sources compiled into oracle/_ref/libref_avatar.so (oracle/ref_optimizer.cpp).
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef AVATAR_ORACLE_H_
#define AVATAR_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;
typedef struct orc_optimizer orc_optimizer;

enum { ORC_SOLVER_BFGS_WOLFE = 0, ORC_SOLVER_GN_LM = 1 };

typedef struct {
    int32_t icp_iters;          /* AvatarOptimizer.h:19 default 1 */
    int32_t max_iters_per_icp;  /* AvatarOptimizer.h:36 default 10 */
    double beta_pose;           /* AvatarOptimizer.h:27 default 0.1 */
    double beta_shape;          /* default 1.0 */
    int32_t enable_occlusion;   /* AvatarOptimizer.h:39 default true */
    int32_t solver;             /* ORC_SOLVER_* */
    double function_tolerance;  /* AvatarOptimizer.cpp:1333 = 1e-4 */
    int32_t num_threads;        /* AvatarOptimizer.h:20 default 4 */
    int32_t nn_method;          /* 0 = brute force, 1 = kd-tree (own), both exact */
} orc_options;

typedef struct {
    int32_t num_correspondences; /* of the last ICP iteration */
    int32_t iterations;          /* solver iterations executed in the last ICP iteration */
    int32_t evaluations;         /* cost/gradient evaluations, all ICP iterations */
    int32_t accepted_steps;
    double initial_cost;         /* last ICP iteration */
    double final_cost;
    double seconds;              /* wall time of the optimize call */
} orc_stats;

/* AvatarModel::AvatarModel npz branch (AvatarModel.cpp:23-127).  Arrays in numpy C order:
 * v_template [V][3], shapedirs [V][3][K], j_regressor [J][V], weights [V][J], parents [J],
 * faces [F][3]. */
orc_model* orc_model_create(int V, int J, int K, int F, const double* v_template,
                            const double* shapedirs, const double* j_regressor,
                            const double* weights, const int32_t* parents, const int32_t* faces);
void orc_model_destroy(orc_model*);
/* GaussianMixture::load maths on already-parsed arrays (GaussianMixture.cpp:20-77).
 * returns 0, or 1 when a covariance is not positive definite ("Decomposition failed!", :60). */
int orc_model_set_prior(orc_model*, int C, int D, const double* weight, const double* mean,
                        const double* cov);
/* GaussianMixture::load from the reference text format; nComps = -1 semantics if unreadable. */
int orc_model_load_prior_text(orc_model*, const char* path);
/* taps */
void orc_model_get_joint_reg(const orc_model*, double* base3J, double* reg3JxK, double* initial3J);
int orc_model_get_assigned(const orc_model*, int32_t* start /*V+1*/, int32_t* joint, double* weight);
void orc_model_get_prior(const orc_model*, double* prec_cho /*C*D*D row-major*/, double* consts_log);

/* Avatar::update (Avatar.cpp:22-75). R: J row-major 3x3. out: cloud 3V (xyz per point),
 * joint_pos 3J, joint_trans 12J (each 3x4 column-major as Eigen stores it). */
void orc_avatar_update(const orc_model*, const double* p, const double* R, const double* w,
                       double* cloud, double* joint_pos, double* joint_trans);
/* AvatarOptimizer.cpp:1250-1254: AngleAxisd::fromRotationMatrix -> Quaterniond (x,y,z,w). */
void orc_rotmat_to_quat(const double* R_rowmajor, double* q_xyzw);
void orc_quat_to_rotmat(const double* q_xyzw, double* R_rowmajor);
/* GaussianMixture::residual (GaussianMixture.cpp:95-114): out has D+1 entries. returns comp. */
int orc_gmm_residual(const orc_model*, const double* x, double* out);

/* AvatarOptimizer::AvatarOptimizer (AvatarOptimizer.cpp:1213-1244) */
orc_optimizer* orc_optimizer_create(const orc_model*, int num_parts, const int32_t* part_map);
void orc_optimizer_destroy(orc_optimizer*);
/* back-face visibility (AvatarOptimizer.cpp:1349-1367); out: V bytes */
void orc_visibility(const orc_optimizer*, const double* cloud, uint8_t* visible);
/* findNN(..., invert=true) (AvatarOptimizer.cpp:841-920): out idx[N] = matched model vertex or -1 */
void orc_find_nn(const orc_optimizer*, const double* cloud, const uint8_t* visible,
                 const double* data /*3N*/, const int32_t* labels, int N, int method, int32_t* idx);
/* RTree::postProcess (RTree.cpp:3422-3450): suppressPartNonMax / removeSmallPieces + upscaleGrid on one label image */
void orc_rtree_postprocess(uint8_t* image, int width, int height, const int32_t* roi, int interval, int num_parts, int part_map_type,
                           double* com_pre, double dist_to_pre_weight);
/* exact brute-force 1-NN with nanoflann's distance arithmetic (no part structure): out[nq] = index into pts */
void orc_brute_nn(const double* pts, int n, const double* queries, int nq, int32_t* out);
/* One evaluation of the Ceres problem (AvatarOptimizer.cpp:283-347, 505-582, 632-639, 661-692,
 * 708-723) at x = [p(3), q(4J xyzw), w(K)] with correspondences idx[N] (-1 = none).
 * beta_* are the UNscaled betaPose/betaShape; scaling by sqrt(#corr)/15 (:1457-1458) is applied.
 * grad: tangent-space gradient (3 + 3J + K); H (nullable): Gauss-Newton J^T J, row-major P x P. */
double orc_evaluate(const orc_optimizer*, const double* x, const double* data,
                    const int32_t* idx, int N, double beta_pose, double beta_shape,
                    int num_threads, double* grad, double* H);
/* per-vertex position and dense tangent Jacobian (3 x P row-major) at x (AvatarOptimizer.cpp:505-582) */
void orc_vertex_jacobian(const orc_optimizer*, const double* x, int vertex, double* pos3,
                         double* jac3xP);
/* independent forward-model restatement following the autodiff functor (AvatarOptimizer.cpp:742-818) */
void orc_vertex_position_chain(const orc_optimizer*, const double* x, int vertex, double* pos3);
/* FakeQuaternionParameterization::Plus etc. (AvatarOptimizer.cpp:123-143): x_plus = x (+) delta */
void orc_retract(const orc_optimizer*, const double* x, const double* delta, double* x_plus);
/* AvatarOptimizer::optimize (AvatarOptimizer.cpp:1246-1517) from quaternions x (in/out).
 * trace (nullable): receives x after every solver iteration of every ICP iteration,
 * capacity trace_cap vectors; *trace_len = number written. nn_out (nullable): idx[N] of the
 * last ICP iteration. */
int orc_optimize(const orc_optimizer*, const double* data, const int32_t* labels, int N,
                 double* x, const orc_options*, orc_stats*, double* trace, int trace_cap,
                 int* trace_len, int32_t* nn_out);
/* Data-cloud construction as the reference's callers do it: CameraIntrin::depthToXYZ
 * (Calibration.cpp:83-95, float arithmetic) followed by the count-and-fill loops of demo.cpp:215-250 over the
 * bounding box roi = {x0, y0, x1, y1} (inclusive; NULL = whole image) at stride `interval`: background = 255,
 * y negated, raster order.  intrin = {fx, cx, fy, cy}.  Returns the number of points, -1 when a label >= num_parts
 * is met (the reference exits, demo.cpp:232-239), -2 when `capacity` points do not suffice. */
int64_t orc_build_cloud(const float* depth, const uint8_t* parts, int width, int height, const float* intrin,
                        const int32_t* roi, int interval, int num_parts, double* cloud, int32_t* labels,
                        int64_t capacity);
/* Image form of RTree::predictBest (RTree.cpp:3184-3262) with upscaleGrid (RTree.cpp:70-100) when fill_in_gaps and
 * interval > 1.  Tree = the arrays of RTree::nodes (u, v [nodes][2], thresh, lnode, rnode, leafid
 * with -1 = internal) and leafBestMatch.  roi = {top_left.x, top_left.y, bot_right.x, bot_right.y} or NULL (whole
 * image).  out: [height][width] uint8, 255 where nothing is predicted.  The reference's row loop pre-increments, so
 * the first row of the box is never predicted; its memset in upscaleGrid may run past bot_right.x (kept) and past the
 * end of the row (clamped to the image here). */
void orc_rtree_predict(const float* depth, int width, int height, int num_nodes, const float* u, const float* v,
                       const float* thresh, const int32_t* lnode, const int32_t* rnode, const int32_t* leafid,
                       const uint8_t* leaf_best, const int32_t* roi, int interval, int fill_in_gaps, uint8_t* out);
/* functional stand-in for AvatarRenderer::renderDepth / renderPartMask used only by tests */
int orc_param_dim(const orc_optimizer*);   /* 3 + 4J + K */
int orc_tangent_dim(const orc_optimizer*); /* 3 + 3J + K */

#ifdef __cplusplus
}
#endif
#endif
