// avb_kernels.h -- kernel argument blocks and launchers (host <-> device interface inside the library)
#pragma once
#include <cuda_runtime.h>
#include "avb_device.cuh"

#define AVB_MAX_ASSIGN_ 4

namespace avb {

struct PoseArgs {
    const double* x;       // [batch][nx]
    double* cloud;         // [batch][3V]
    double* joint_pos;     // nullable [batch][3J]
    double* joint_trans;   // nullable [batch][12J]
    int do_visibility;     // 0: forward only (final ava.update())
    int do_lbs;            // 0: the cloud is already posed (a sliced forward-only launch ran before): visibility + compaction only
    int slices;            // forward-only launches of small batches: CTAs per frame, each poses a slice of the vertices (grid = batch * slices)
    int enable_occlusion;
    uint8_t* visible;      // [batch][V]
    int* pv_idx;           // [batch][V]       compacted (part, vertex id) order -> vertex id
    double* pv_xyz;        // [batch][pv_stride] compacted positions
    int* pv_start;         // [batch][numParts+1]
    long long pv_stride;   // doubles per frame (even, >= 3V + 2)
    float4* pv_f32;        // nullable [batch][V]: the compacted positions relative to the root translation p, as floats (nn_kernel's pre-scan)
    float* pv_rmax;        // [batch]: max |coordinate| of pv_f32 of the frame (error bound of the pre-scan)
    int* zero_cnt;                   // nullable [batch][V]: the correspondence counters nn_kernel adds into, zeroed here with the
    unsigned long long* zero_sum;    // [batch][V][3] sums and the frame's range flag (instead of three memset launches per fit)
    int* zero_range;                 // [batch]
};

struct NNArgs {
    int V;
    const double* data;          // [3 * total points]
    const float* data_f32;       // nullable: the same cloud as uploaded floats (avb_upload_batch_f32); widened on load, bit-identical
    const int* labels;           // [total points]
    const int* chunk_frame;      // [chunks]
    const long long* chunk_begin;  // [chunks] first global point of the chunk
    const int* chunk_count;      // [chunks]
    const int* chunk_qblock;     // [chunks] index of the chunk's first |d|^2 partial
    const int* pv_idx;
    const double* pv_xyz;
    const int* pv_start;
    long long pv_stride;
    const float4* pv_f32;        // nullable: float copy relative to x[f][0..2] (pose_visibility_kernel); null = plain fp64 scan
    const float* pv_rmax;        // [batch]
    const double* x;             // [batch][nx] the parameters the cloud was posed with (root translation = centre of pv_f32)
    int nx;
    int* nn_idx;                 // [total points]
    int* cnt;                    // [batch][V]
    unsigned long long* sum;     // [batch][V][3] fixed point 2^36
    double* qpart;               // [total q blocks]
    int* range_flag;             // [batch]
    int stage_cap;               // compacted model vertices staged in shared memory (the rest is read through L2)
};

struct CloudArgs {  // data-cloud construction from depth + part-label images (avb_cloud.cu)
    const float* depth;          // [batch][height][width] metres
    const uint8_t* parts;        // [batch][height][width] body-part label, 255 = background
    const int* roi;              // nullable [batch][4] x0, y0, x1, y1 inclusive (demo.cpp: bgsub.topLeft / botRight)
    int width, height, interval, num_parts;
    float fx, cx, fy, cy;        // CameraIntrin (Calibration.cpp:68-74)
    int max_strips;              // row stride of strip_count / strip_offset
    int* strip_count;            // [batch][max_strips] foreground pixels per strip of 8 sampled rows
    const long long* strip_offset;  // [batch][max_strips] first point of the strip in the batch cloud
    int* bad_label;              // [batch] set when a label >= num_parts (and != 255) was seen
    double* cloud;               // [3 * total points] out
    int* labels;                 // [total points] out
};

struct RTreeNode {   // RTree::RNode (include/RTree.h:28-41) packed into one 32-byte sector
    float ux, uy, vx, vy, thresh;
    int lnode, rnode, leafid;   // leafid == -1: internal node
};
struct RTreeArgs {   // RTree::predictBest on images (avb_rtree.cu)
    const RTreeNode* nodes;
    const uint8_t* leaf_best;    // [leaves] leafBestMatch
    const float* depth;          // [batch][height][width]
    uint8_t* parts;              // [batch][height][width] out (255 = not predicted)
    const int* roi;              // nullable [batch][4] top_left.x, top_left.y, bot_right.x, bot_right.y
    int width, height, interval;
};

struct RTreePostArgs {   // RTree::postProcess on the device (avb_rtree.cu)
    uint8_t* parts;              // [batch][height][width] in/out
    const int* roi;              // nullable [batch][4]
    int width, height, interval, num_parts, part_map_type;
    double dist_w;               // dist_to_pre_weight
    double* com_pre;             // [batch][2 * num_parts] in/out, x of part i at [2 i] (-1: not seen), y at [2 i + 1]
    int* arena;                  // [batch][cap] component pixel lists (append-only per frame)
    int* stack;                  // [batch][cap] DFS stack
    long long cap;               // ints per frame in arena and in stack (>= 2 * grid pixels + 64)
    int* overflow;               // [batch] set when a frame ran out of scratch
};

struct RenderArgs {   // AvatarRenderer on the device (avb_render.cu)
    const double* cloud;         // [batch][3V] posed models
    const int* faces;            // [3F]
    const uint8_t* vpart;        // [V] part of the vertex's main joint
    float* proj;                 // [batch][V][2] scratch: projected vertices
    int* order;                  // [batch][F] scratch: faces in paint order
    unsigned* win_depth;         // nullable [batch][H][W] scratch: winning rank per pixel, zeroed by the caller
    unsigned* win_parts;
    unsigned* win_faces;
    float* depth_out;            // nullable [batch][H][W]
    uint8_t* parts_out;
    int* faces_out;
    int V, F, width, height;
    float fx, cx, fy, cy;
    // renderLambert (AvatarRenderer.cpp:103-172)
    const int* vf_start;         // [V+1] CSR of the faces incident to every vertex
    const int* vf_list;          // [3F]
    int* rank_of;                // nullable [batch][F] scratch: paint position of every model face
    float* vlam;                 // [batch][V] scratch: the value painted at every vertex
    unsigned* win_lambert;       // [batch][H][W] scratch, zeroed by the caller
    uint8_t* lambert_out;        // [batch][H][W]
};

struct LmState {  // per-frame Levenberg-Marquardt state, lives in HBM between the kernels of one ICP iteration
    double cost, radius, decrease, Qsum, sbp, sbs, initial_cost, model_change;
    int done, iters, accepted, ncorr, nmatched, nchunks, evals, nslots;   // nslots: record slots incl. alignment gaps
    int last, pad0, pad1, pad2;   // last: the pending evaluation only decides the final accept / reject (cost only)
};

struct FlowQueue {  // device work queue of lm_flow_kernel
    unsigned long long* slots;   // [cap_mask + 1] ring of (ticket << 32 | task); nullptr: staged kernels
    unsigned int* ctrl;          // [0] head, [1] tail, [2] frames still running, [3] watchdog flag, [4] CTAs that left
    int* rows_left;              // [batch] record blocks of the current evaluation still running
    int* gram_left;              // [batch] chunks of the current evaluation still running
    unsigned long long* prof;    // nullable [16]: CTA nanoseconds in rows / gram / solve tasks / waiting, then sub-phases
    unsigned int cap_mask;
};

struct LmBuf {
    double* x;                   // [batch][nx] current point (in/out)
    double* xt;                  // [batch][nx] trial point
    double* tab;                 // [batch][tabD] joint tables of the trial point: G | pos | tau | C
    unsigned short* mlist;       // [batch][rec_rs] matched vertices grouped by Jacobian column group (0xFFFF = gap)
    int4* chunks;                // [batch][maxc] (group, start in mlist, count, -)
    int2* gruns;                 // [batch][kMaxGroups] (first chunk, chunks) of every column group
    double* part;                // [batch][maxc][pstride] chunk partial: upper triangle of J^T J | J^T r, group columns
    double* cpart;               // [batch][maxrb] cost partial per 256 record slots
    float* rec;                  // [batch][rec_stride][rec_rs] fp32 Jacobian records (SoA) of the matched vertices
    int* gstart;                 // [batch][kMaxGroups+1] group boundaries inside mlist
    int maxrb, rec_stride;       // rec_stride: fields per record (max over groups)
    int rec_rs;                  // record slots per frame (multiple of 4)
    double* gcur;                // [batch][P]
    double* Hcur;                // [batch][P*P]
    LmState* state;              // [batch]
    int maxc, tabD, chunk_verts;
    long long pstride;
    const int* cnt;              // [batch][V]
    const unsigned long long* sum;  // [batch][V][3] fixed point 2^36
    const double* qpart;
    const int* frame_qblock;     // [batch+1]
    const int* range_flag;       // [batch]
    double beta_pose, beta_shape, function_tolerance;
    int max_iters;
    double* dump_cost;           // nullable [batch]  (avb_debug_evaluate)
    double* dump_grad;           // [batch][P]
    double* dump_H;              // [batch][P*P]
    double* trace;               // nullable [batch][trace_cap][nx]
    int trace_cap;
    FrameStats* stats;           // [batch]
    int tensor;                  // 1: fused record + Gram tasks on tcgen05 (lm_flow_kernel<true>); 0: fp64 DMMA path
    int fused;                   // fp64 path of lm_flow_kernel: 1 = one fused record + Gram task per chunk (no d_rec round trip)
    int prior_task;              // lm_flow_kernel, small batches: the pose prior of a trial point is its own task, evaluated next to the
                                 // record / Gram tasks on another SM (same arithmetic; the solve only adds the result)
    double* prior_out;           // [batch][prior_stride]: prior cost | winning component | y = Sigma^-1 (x - mu) of that component
    int prior_stride;
    FlowQueue q;
};

cudaError_t launch_rtree_predict(const RTreeArgs& a, int batch, int max_box_pixels, cudaStream_t st);
cudaError_t launch_rtree_upscale(const RTreeArgs& a, int batch, int max_box_pixels, cudaStream_t st);
cudaError_t launch_rtree_postprocess(const RTreePostArgs& a, int batch, cudaStream_t st);
int render_max_faces();
cudaError_t launch_render(const RenderArgs& a, int batch, cudaStream_t st, cudaEvent_t* ev4);
int render_max_valence();
cudaError_t launch_render_lambert(const RenderArgs& a, int batch, cudaStream_t st);
int cloud_strip_rows();
cudaError_t launch_cloud_count(const CloudArgs& a, int strips, int batch, cudaStream_t st);
cudaError_t launch_cloud_compact(const CloudArgs& a, int strips, int batch, cudaStream_t st);
size_t pose_smem_bytes(int V, int J, int K);
size_t nn_smem_bytes(int stage_cap, bool f32);
cudaError_t launch_widen_points(const float* in, double* out, long long n, int num_sms, cudaStream_t st);
cudaError_t launch_pose_visibility(const DevModel& M, const DevParts& Pt, const PoseArgs& a, int batch, cudaStream_t st);
cudaError_t launch_nn(const DevParts& Pt, const NNArgs& a, int num_chunks, cudaStream_t st);
cudaError_t launch_lm_prep(const DevModel& M, const DevParts& Pt, const LmBuf& a, int batch, cudaStream_t st);
// one evaluation = parts 0 (lm_rows_kernel), 1 (lm_gram_kernel or lm_gram_tc_kernel), 2 (lm_solve_kernel) in order
cudaError_t launch_lm_eval_part(const DevModel& M, const DevParts& Pt, const LmBuf& a, int batch, int max_nj, bool tensor,
                                int part, cudaStream_t st);
// the same evaluations as one persistent data-flow kernel (launch_lm_prep with a.q set seeds the queue)
cudaError_t launch_lm_flow(const DevModel& M, const DevParts& Pt, const LmBuf& a, int max_nj, int ctas, int occ, cudaStream_t st);
size_t lm_flow_smem_bytes(const DevModel& M, int max_nj, int chunk_verts, bool tensor);
long long lm_part_stride(int max_nj, int J, int K);
int lm_tab_doubles(int J, int K);
int lm_rec_floats(int max_nj, int K);
int lm_rec_slots(int V);
size_t lm_gram_smem_bytes(int max_nj, int K, int chunk_verts, bool tensor);
bool lm_tensor_supported(int max_nj, int K);

}  // namespace avb
