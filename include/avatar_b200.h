/* avatar_b200.h -- C ABI of the B200-native SMPL-to-point-cloud fitting engine.
 *
 * Drop-in boundary for ONE path of sxyu/avatar: ark::AvatarOptimizer::optimize and what it
 * calls (Avatar::update, GaussianMixture, the nanoflann correspondence search, the Ceres solve).
 * The reference has no FFI; its boundary is the C++ class API of include/Avatar.h and
 * include/AvatarOptimizer.h.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  The header-compatible C++ facade
 * (include/ark_b200/, avatar_b200/cpp/ark_b200.cpp) and the Python mirror (the avatar_b200 package) are thin callers of this ABI.
 *
 * Conventions: int return codes (AVB_OK = 0), no exceptions cross the ABI, caller owns host
 * buffers, the library owns device buffers, plain pointers and sizes only.  All host arrays are
 * fp64 / int32 exactly as the reference's Eigen types hold them.
 *
 * There is NO CPU fallback: every compute entry point fails with AVB_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef AVATAR_B200_H_
#define AVATAR_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AVB_VERSION 100
#define AVB_MAX_ASSIGN 4 /* AvatarEvaluationCommonData::MAX_ASSIGN, AvatarOptimizer.cpp:164 */
#define AVB_MAX_JOINTS 32
#define AVB_MAX_SHAPE_KEYS 16

enum {
    AVB_OK = 0,
    AVB_ERR_INVALID = 1,   /* bad argument (null pointer, sizes, labels out of range, ...) */
    AVB_ERR_CUDA = 2,      /* CUDA runtime / driver error, or no usable device */
    AVB_ERR_CAPACITY = 3,  /* batch / point count exceeds what the fitter was created for */
    AVB_ERR_NUMERIC = 4,   /* non-finite input, coordinate out of range (|x| >= 32 m), GMM not PD */
    AVB_ERR_PRIOR = 5      /* betaPose > 0 but the model has no pose prior (the reference would
                              dereference an empty GMM here, AvatarOptimizer.cpp:1465) */
};

typedef struct avb_model avb_model;   /* ~ ark::AvatarModel (immutable, shareable) */
typedef struct avb_fitter avb_fitter; /* ~ ark::AvatarOptimizer (+ the device side of ark::Avatar) */

/* Host description of an ark::AvatarModel (include/Avatar.h:64-151), in the reference's own
 * member layouts.  num_points=V, num_joints=J, num_shape_keys=K, num_faces=F. */
typedef struct {
    int32_t num_points, num_joints, num_shape_keys, num_faces;
    const double* base_cloud;           /* AvatarModel::baseCloud            [3V]  x0 y0 z0 x1 ... */
    const double* key_clouds;           /* AvatarModel::keyClouds  (3V x K), row 3v+c, ROW-major [3V][K] */
    const double* joint_shape_reg_base; /* AvatarModel::jointShapeRegBase    [3J] */
    const double* joint_shape_reg;      /* AvatarModel::jointShapeReg (3J x K), ROW-major [3J][K] */
    const int32_t* parent;              /* AvatarModel::parent               [J], parent[0] = -1 */
    const int32_t* mesh;                /* AvatarModel::mesh (3 x F col-major) = [F][3] vertex ids */
    /* AvatarModel::assignedJoints as CSR: per vertex (weight, joint) sorted descending
     * (AvatarModel.cpp:74-94); at most AVB_MAX_ASSIGN entries per vertex. */
    const int32_t* assign_start;        /* [V+1] */
    const int32_t* assign_joint;        /* [assign_start[V]] */
    const double* assign_weight;        /* [assign_start[V]] */
    /* GaussianMixture posePrior as parsed from pose_prior.txt (GaussianMixture.cpp:20-58);
     * gmm_components <= 0 means "no prior" (hasPosePrior() == false). gmm_dims must be 3(J-1). */
    int32_t gmm_components, gmm_dims;
    const double* gmm_weight;           /* [C] */
    const double* gmm_mean;             /* [C][D] */
    const double* gmm_cov;              /* [C][D][D] */
} avb_model_desc;

typedef struct {
    int32_t device;             /* CUDA device ordinal */
    int32_t max_batch;          /* frames per avb_fit_batch call */
    int64_t max_total_points;   /* sum of N over a batch */
    int32_t num_parts;          /* AvatarOptimizer ctor: num_parts */
    const int32_t* part_map;    /* AvatarOptimizer ctor: part_map [J] joint -> part */
} avb_fitter_config;

enum { AVB_SOLVER_GN_LM = 1 };
/* precision of the J^T J accumulation: FP64 (default; fitted parameters track the fp64 oracle to ~1e-6),
 * FP32 (faster; ~1e-4 drift over 10 iterations), BF16_TENSOR (tcgen05 path, BASELINE.json configs[4]) */
enum { AVB_JTJ_FP64 = 0, AVB_JTJ_FP32 = 1, AVB_JTJ_BF16_TENSOR = 2 };

/* Public tunables of ark::AvatarOptimizer (include/AvatarOptimizer.h:27-39) plus the solver
 * controls the reference hard-codes in optimize() (AvatarOptimizer.cpp:1313-1341). */
typedef struct {
    int32_t icp_iters;          /* optimize(..., icp_iters = 1, ...) */
    int32_t max_iters_per_icp;  /* maxItersPerICP = 10 */
    double beta_pose;           /* betaPose = 0.1 */
    double beta_shape;          /* betaShape = 1.0 */
    int32_t enable_occlusion;   /* enableOcclusion = true */
    int32_t nn_step;            /* nnStep = 20: accepted and ignored -- it only affects the
                                   reference's unused forward-NN mode (AvatarOptimizer.cpp:933) */
    double function_tolerance;  /* options.function_tolerance = 1e-4 (:1333) */
    int32_t solver;             /* AVB_SOLVER_GN_LM (BASELINE.json north_star) */
    int32_t jtj_precision;      /* AVB_JTJ_* */
    int32_t reserved[4];
} avb_options;

/* Per-frame fit statistics. */
typedef struct {
    int32_t num_points;          /* N of this frame */
    int32_t num_correspondences; /* data points matched in the last ICP iteration */
    int32_t num_matched_vertices;
    int32_t iterations;          /* LM iterations executed in the last ICP iteration */
    int32_t accepted_steps;
    int32_t status;              /* AVB_OK or AVB_ERR_NUMERIC */
    double initial_cost;         /* last ICP iteration */
    double final_cost;
} avb_stats;

void avb_default_options(avb_options* out);
const char* avb_last_error(void); /* thread-local message of the last failing call */
int avb_device_count(void);

/* replaces: ark::AvatarModel::AvatarModel (AvatarModel.cpp:14-298) -- the data, not the file I/O;
 * the GMM maths of GaussianMixture::load (GaussianMixture.cpp:22-76) runs here. */
int avb_model_create(const avb_model_desc* desc, avb_model** out);
void avb_model_destroy(avb_model* model);
int avb_model_dims(const avb_model* model, int32_t* V, int32_t* J, int32_t* K, int32_t* F);
/* taps for tests: prec_cho [C][D][D] (row-major, lower), consts_log [C] as GaussianMixture holds them */
int avb_model_get_prior(const avb_model* model, double* prec_cho, double* consts_log);

/* replaces: ark::AvatarOptimizer::AvatarOptimizer (AvatarOptimizer.cpp:1213-1244) */
int avb_fitter_create(const avb_model* model, const avb_fitter_config* cfg, avb_fitter** out);
void avb_fitter_destroy(avb_fitter* fitter);

/* Parameter vector of one frame: x = [ p(3) | q_0..q_{J-1} (x,y,z,w each) | w(K) ], 3+4J+K doubles:
 * ark::Avatar::p, AvatarOptimizer::r (Eigen::Quaterniond coeffs order), ark::Avatar::w. */
int avb_param_dim(const avb_model* model);   /* 3 + 4J + K */
int avb_tangent_dim(const avb_model* model); /* 3 + 3J + K */

/* replaces: the R -> AngleAxis -> Quaternion prologue of optimize() (AvatarOptimizer.cpp:1250-1254)
 * and its inverse at :1494-1496.  Host-side, row-major 3x3. */
void avb_rotmat_to_quat(const double* R_rowmajor, double* q_xyzw);
void avb_quat_to_rotmat(const double* q_xyzw, double* R_rowmajor);

/* replaces: ark::Avatar::update (Avatar.cpp:22-75) for `batch` parameter vectors.
 * Outputs (each nullable): cloud [batch][3V], joint_pos [batch][3J], joint_trans [batch][12J]
 * (3x4 column-major per joint, as Avatar::jointTrans). */
int avb_avatar_update(avb_fitter* fitter, int batch, const double* x, double* cloud,
                      double* joint_pos, double* joint_trans);

/* replaces: ark::AvatarOptimizer::optimize (AvatarOptimizer.cpp:1246-1517) for one frame.
 * data_cloud: 3 x N column-major (x0 y0 z0 x1 ...), data_part_labels: [N] in [0, num_parts).
 * x is read as the warm start and overwritten with the fit. cloud_out (nullable): ava.cloud [3V]
 * after the trailing ava.update() (:1497). */
int avb_fit(avb_fitter* fitter, const double* data_cloud, const int32_t* data_part_labels, int32_t n,
            double* x, const avb_options* opt, avb_stats* stats, double* cloud_out);

/* `batch` independent frames in one call (BASELINE.json configs[2]): clouds concatenated,
 * frame f owns points [offsets[f], offsets[f+1]).  x: [batch][3+4J+K] in/out; stats: [batch]
 * (nullable); cloud_out: [batch][3V] (nullable).  Host buffers; for peak transfer speed allocate
 * them with avb_host_alloc (pinned). */
int avb_fit_batch(avb_fitter* fitter, int32_t batch, const double* data_clouds,
                  const int32_t* data_part_labels, const int64_t* offsets, double* x,
                  const avb_options* opt, avb_stats* stats, double* cloud_out);

/* Tracking mode (BASELINE.json configs[3]): ONE sequence of T frames fitted in order, frame t warm-started from the
 * fit of frame t-1 -- what the reference's callers do by keeping `ava` between frames (demo.cpp:252-268,
 * live-demo.cpp:405-432).  All clouds are passed at once (concatenated, offsets[T+1]); their H2D copies run ahead on
 * a copy stream (one event per frame) while earlier frames are being fitted, and the parameter vector never leaves
 * the device between frames.  x0: start of frame 0 [nx]; x_out: [T][nx]; stats: [T] (nullable). */
int avb_track_sequence(avb_fitter* fitter, int32_t T, const double* data_clouds, const int32_t* data_part_labels,
                       const int64_t* offsets, const double* x0, const avb_options* opt, double* x_out,
                       avb_stats* stats);

/* Split form of avb_fit_batch for callers that keep inputs resident on the device
 * (bench.py `value`): upload once, fit many times from different warm starts, download. */
int avb_upload_batch(avb_fitter* fitter, int32_t batch, const double* data_clouds,
                     const int32_t* data_part_labels, const int64_t* offsets);
/* The same with the cloud as FLOATS (12 instead of 24 bytes per point over PCIe).  Depth cameras deliver float points
 * (CameraIntrin::to3D / depthToXYZ compute in float, Calibration.cpp:68-95) and the reference widens them on the host
 * when it fills the Eigen cloud (demo.cpp:241-243); here the widening runs on the device and yields the same doubles,
 * so every result is bit-identical to avb_upload_batch of the widened cloud. */
int avb_upload_batch_f32(avb_fitter* fitter, int32_t batch, const float* data_clouds,
                         const int32_t* data_part_labels, const int64_t* offsets);
int avb_fit_resident(avb_fitter* fitter, const double* x_in, const avb_options* opt);  /* async enqueue */

/* Data-cloud construction on the device.  Replaces the caller-side loops of
 * demo.cpp:215-250 / live-demo.cpp:364-400 (count and fill dataCloud / dataPartLabels from the xyz map and the
 * RTree label image inside the background-subtractor bounding box, stride `interval`, y negated) together with
 * CameraIntrin::depthToXYZ (Calibration.cpp:83-95, float arithmetic).  depth: [batch][height][width] float metres;
 * parts: [batch][height][width] uint8, 255 = background; roi: nullable [batch][4] = x0, y0, x1, y1 inclusive
 * (bgsub.topLeft / botRight; NULL = whole image).  Points are produced in the reference's raster order and are
 * bit-identical to the host loop.  A label >= num_parts other than 255 is an error (the reference prints FATAL and
 * exits, demo.cpp:232-239).  Afterwards the batch is resident exactly as after avb_upload_batch: call
 * avb_fit_resident / avb_download_results; offsets_out (nullable, batch + 1 entries) receives the per-frame point
 * offsets; avb_download_batch reads the constructed clouds back (tests).
 * parts == NULL: the labels are predicted on the device with the decision tree given to avb_fitter_set_rtree, as
 * demo.cpp:198-200 does on the host (RTree::predictBest(depth, threads, rtree_interval, topLeft, botRight) with gap
 * filling), so only the depth image crosses PCIe. */
typedef struct avb_image_desc {
    int32_t width, height;
    float fx, cx, fy, cy;   /* CameraIntrin, in the order of its intrin[] file format (Calibration.cpp:68-74) */
    int32_t interval;       /* pixel stride in both directions (demo.cpp `interval`) */
    int32_t num_parts;      /* rtree.numParts */
    int32_t rtree_interval; /* parts == NULL only: stride of RTree::predictBest (demo.cpp:198 uses 2); gaps are filled */
    int32_t rtree_postprocess;  /* parts == NULL only: != 0 runs RTree::postProcess on the predicted labels (demo.cpp:201), with
                                 * the fitter's per-frame centre-of-mass state (comPre), see avb_rtree_reset_tracking */
    int32_t part_map_type;      /* RTree::partMapType of the .partmap file: 0 = contiguous parts (best blob per part), else
                                 * small pieces are removed (RTree.cpp:3437-3446) */
    double dist_to_pre_weight;  /* postProcess' dist_to_pre_weight; <= 0 selects the reference default 0.001 */
} avb_image_desc;
int avb_upload_depth_batch(avb_fitter* fitter, int32_t batch, const float* depth, const uint8_t* parts,
                           const int32_t* roi, const avb_image_desc* img, int64_t* offsets_out);
int avb_download_batch(avb_fitter* fitter, double* data_clouds, int32_t* data_part_labels, int64_t* offsets);
/* The reference's renderer on the device: AvatarRenderer::renderDepth,
 * renderPartMask and renderFaces (AvatarRenderer.cpp:72-98, 170-216; painters in AvatarHelpers.cpp:61-302) for a batch
 * of parameter vectors x [batch][nx] (the fitter poses the model, then paints).  Outputs are host arrays
 * [batch][height][width], each nullable: depth float (0 = nothing), parts uint8 (255 = nothing, part of the nearest
 * projected vertex otherwise, with the fitter's part_map), faces int32 (-1 = nothing, else the face's position in paint
 * order, as the reference writes it).  Bit-identical to the sequential painter; faces with equal depth keys are painted
 * in ascending face index (the reference's std::sort leaves that order unspecified).  renderLambert: avb_render_lambert_batch.
 * The call poses the models into the fitter's model-cloud buffer: download the results of a pending fit first. */
typedef struct avb_render_desc {
    int32_t width, height;
    float fx, cx, fy, cy;
} avb_render_desc;
int avb_render_batch(avb_fitter* fitter, int32_t batch, const double* x, const avb_render_desc* view, float* depth_out,
                     uint8_t* parts_out, int32_t* faces_out);
/* AvatarRenderer::renderLambert (AvatarRenderer.cpp:103-172), the call demo.cpp:275-277 and live-demo.cpp:428-434 make right
 * after optimize(): per-vertex normals (unit face normals added in the frame's paint order, normalised, turned towards
 * the camera), two point lights, faces with |n_z| <= 1e-2 skipped, painted far to near with paintTriangleBary<uint8_t>.
 * gray_out: host [batch][height][width] uint8 (0 = nothing painted), bit for bit the reference's image. */
int avb_render_lambert_batch(avb_fitter* fitter, int32_t batch, const double* x, const avb_render_desc* desc, uint8_t* gray_out);

/* device time (ms) of [prepare (project + sort), cover, resolve] of the last avb_render_batch */
int avb_last_render_ms(avb_fitter* fitter, float* ms3);
/* Body-part label prediction on the device: the image form of RTree::predictBest
 * (RTree.cpp:3184-3262) with its gap filling (upscaleGrid, RTree.cpp:70-100).  The tree is given as the arrays of
 * RTree::nodes (include/RTree.h:28-41: u, v [nodes][2], thresh, lnode, rnode, leafid with -1 = internal) and
 * leafBestMatch ([leaves], RTree.cpp:3455-3463); file parsing (RTree::loadFile) stays with the caller.
 * avb_rtree_predict_batch: depth [batch][height][width] float (0 = no data) in, labels [batch][height][width] uint8
 * out (255 = not predicted), roi = top_left / bot_right per frame (NULL = whole image), labels bit-exact.  The
 * reference never predicts the first row of the box (its row loop pre-increments, RTree.cpp:3196-3199); neither does
 * this.  postProcess (connected components, RTree.cpp:3422-3450) is not part of it. */
typedef struct avb_rtree_desc {
    int32_t num_nodes, num_leaves, num_parts;
    const float* u;            /* [num_nodes][2] */
    const float* v;            /* [num_nodes][2] */
    const float* thresh;       /* [num_nodes] */
    const int32_t* lnode;      /* [num_nodes] */
    const int32_t* rnode;      /* [num_nodes] */
    const int32_t* leafid;     /* [num_nodes], -1 = internal node */
    const uint8_t* leaf_best;  /* [num_leaves] */
} avb_rtree_desc;
int avb_fitter_set_rtree(avb_fitter* fitter, const avb_rtree_desc* tree);
int avb_rtree_predict_batch(avb_fitter* fitter, int32_t batch, const float* depth, int32_t width, int32_t height,
                            const int32_t* roi, int32_t interval, int32_t fill_in_gaps, uint8_t* parts_out);
/* RTree::postProcess (RTree.cpp:3422-3450; suppressPartNonMax :125-232 or removeSmallPieces :234-320, then upscaleGrid
 * :70-100) on a batch of label images: parts host [batch][height][width] uint8 in/out (255 = background), roi nullable
 * [batch][4] inclusive corners inside the image, com_pre host [batch][2 * num_parts] in/out laid out like the
 * reference's 2 x numParts matrix (x of part i at [2 i], -1 = not seen; y at [2 i + 1]); a fresh matrix is (-1, 0).
 * Bit for bit the reference's sequential flood fill, including its behaviour for interval > 1 (the downward probe tests
 * row r + 1 but continues at row r + interval): one warp per frame walks the image in the reference's order. */
int avb_rtree_postprocess_batch(avb_fitter* fitter, int32_t batch, uint8_t* parts, int32_t width, int32_t height, const int32_t* roi,
                                int32_t interval, int32_t num_parts, int32_t part_map_type, double* com_pre, double dist_to_pre_weight);
/* forget the centres of mass the chained pipeline (avb_upload_depth_batch with parts == NULL and rtree_postprocess) keeps
 * per batch slot between calls: the next call starts from the reference's fresh comPre */
int avb_rtree_reset_tracking(avb_fitter* fitter);

/* device time (ms) of the last RTree prediction (predict + gap filling kernels) */
int avb_last_rtree_ms(avb_fitter* fitter, float* ms);
/* Multi-GPU: frames are independent, so a batch shards as one block of frames per rank (one process
 * or thread per GPU, one fitter each) with NO collective on the data path, and ONE all-gather of the fitted parameters at the
 * end.  These three entry points are that gather for callers in the reference's host language (no Python, no
 * torch.distributed): NCCL is loaded at run time (libnccl.so.2).
 *   avb_comm_unique_id    rank 0 creates the 128-byte NCCL id and hands it to the other ranks by its own means
 *   avb_fitter_comm_init  every rank: ncclCommInitRank on the fitter's device
 *   avb_gather_params     ncclAllGather straight from the device parameter block of the last fit (no host staging of the
 *                         input, no allocation); all_x is host [nranks][max_batch][nx] -- rank r's frames are rows
 *                         [r * max_batch, r * max_batch + its batch)
 *   avb_gather_params_begin / _end   the same in two halves: begin snapshots the parameter block and enqueues the collective
 *                         on a stream of its own and returns at once, so the caller can upload and fit the next batch while
 *                         the collective waits for the peers (and for an SM: a persistent fit kernel of another fitter on the
 *                         same GPU may hold every slot for milliseconds); end waits and copies out.  They must alternate. */
int avb_comm_unique_id(uint8_t* id128);
int avb_fitter_comm_init(avb_fitter* fitter, const uint8_t* id128, int32_t rank, int32_t nranks);
int avb_gather_params(avb_fitter* fitter, double* all_x);
int avb_gather_params_begin(avb_fitter* fitter);
int avb_gather_params_end(avb_fitter* fitter, double* all_x);

/* device time (ms) of [cloud_count_kernel, cloud_compact_kernel] of the last avb_upload_depth_batch (CUDA events) */
int avb_last_cloud_ms(avb_fitter* fitter, float* ms2);
int avb_download_results(avb_fitter* fitter, double* x_out, avb_stats* stats, double* cloud_out);
int avb_synchronize(avb_fitter* fitter);
/* device time of the kernels enqueued by the last avb_fit_resident, measured with CUDA events on the
 * fitter's stream: total over all ICP iterations, and the per-kernel breakdown of the LAST ICP iteration
 * [pose_visibility, nn (with its memsets), lm, final_pose] in ms (nullable). */
int avb_last_device_ms(avb_fitter* fitter, float* total_ms, float* per_kernel_ms4);
/* CUDA-event stopwatch on the fitter's stream (bench.py): start records an event, stop records another,
 * synchronises and returns the elapsed device time between them. */
int avb_timer_start(avb_fitter* fitter);
int avb_timer_stop(avb_fitter* fitter, float* ms);
/* Per-kernel timing for bench.py's roofline: when enabled, the next avb_fit_resident records a CUDA-event pair
 * around every kernel launch; avb_last_kernel_ms then returns, per kernel class
 * [pose_visibility, nn, lm_prep, lm_rows, lm_gram, lm_solve, final_pose, lm_flow], the summed device time (ms) and
 * the number of launches.  Costs ~2 events per launch, so leave it off inside throughput-timed regions.
 * The inner solve normally runs as ONE persistent kernel (lm_flow_kernel: record, Gram and solve tasks of all frames
 * from a device work queue), so lm_rows / lm_gram / lm_solve are zero unless the staged kernels are selected
 * (AVB_JTJ_BF16_TENSOR, or AVB_FLOW=0 in the environment); avb_last_flow_task_ms returns the CTA time (ms, summed
 * over CTAs) the profiled lm_flow_kernel spent in [record tasks, Gram tasks, solves, waiting for work]. */
int avb_set_profiling(avb_fitter* fitter, int enabled);
int avb_last_kernel_ms(avb_fitter* fitter, float* total_ms8, int32_t* launches8);
int avb_last_flow_task_ms(avb_fitter* fitter, float* ms4);
/* Finer split of the same profiled launch (CTA ms): solve [load, partial reduction, basis change, pose prior,
 * step control, Cholesky, back substitution, retraction + tables], record task [prologue, body], Gram task
 * [tile load, DMMA loop] (the Gram epilogue is the rest of the Gram task time). */
int avb_last_flow_phase_ms(avb_fitter* fitter, float* ms12);
/* The fitter's static Jacobian column groups (vertices whose skinning joints share an ancestor set): number of
 * groups, joints per group and model vertices per group (arrays of 16).  bench.py derives the algorithmic bytes
 * of the record / Gram kernels from them. */
int avb_fitter_groups(avb_fitter* fitter, int32_t* num_groups, int32_t* joints16, int32_t* vertices16);
/* number of kernel launches enqueued by the last avb_fit_resident / avb_fit_batch */
int avb_last_launch_count(avb_fitter* fitter);

/* pinned host memory for the host-side batches (cudaHostAlloc) */
void* avb_host_alloc(uint64_t bytes);
void avb_host_free(void* p);

/* Parity taps (tests only; they run the same kernels the fit runs, on the uploaded batch):
 *   AVB_TAP_VISIBLE   uint8  [batch][V]   back-face visibility (AvatarOptimizer.cpp:1349-1367)
 *   AVB_TAP_NN        int32  [total N]    matched model vertex per data point, -1 = none (:841-920)
 *   AVB_TAP_CLOUD     double [batch][3V]  model cloud the NN pass searched
 *   AVB_TAP_COUNT     int32  [batch][V]   #data points matched to each vertex
 *   AVB_TAP_SUM       double [batch][3V]  sum of the data points matched to each vertex */
enum { AVB_TAP_VISIBLE = 1, AVB_TAP_NN = 2, AVB_TAP_CLOUD = 3, AVB_TAP_COUNT = 4, AVB_TAP_SUM = 5 };
/* run pose+visibility+NN once at x_in on the uploaded batch (no solve) so the taps can be read */
int avb_debug_correspond(avb_fitter* fitter, const double* x_in, const avb_options* opt);
int avb_debug_read(avb_fitter* fitter, int what, void* out, uint64_t bytes);
/* one evaluation of the objective at x_in with the correspondences currently on the device:
 * cost [batch], grad [batch][P] (tangent space), H [batch][P*P] (Gauss-Newton matrix incl. priors).
 * Mirrors one Ceres evaluation (AvatarOptimizer.cpp:283-347, 505-582, 632-639, 661-692, 708-723). */
int avb_debug_evaluate(avb_fitter* fitter, const double* x_in, const avb_options* opt, double* cost,
                       double* grad, double* H);

#ifdef __cplusplus
}
#endif
#endif /* AVATAR_B200_H_ */
