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
    int enable_occlusion;
    uint8_t* visible;      // [batch][V]
    int* pv_idx;           // [batch][V]       compacted (part, vertex id) order -> vertex id
    double* pv_xyz;        // [batch][pv_stride] compacted positions
    int* pv_start;         // [batch][numParts+1]
    long long pv_stride;   // doubles per frame (even, >= 3V + 2)
};

struct NNArgs {
    int V;
    const double* data;          // [3 * total points]
    const int* labels;           // [total points]
    const int* chunk_frame;      // [chunks]
    const long long* chunk_begin;  // [chunks] first global point of the chunk
    const int* chunk_count;      // [chunks]
    const int* chunk_qblock;     // [chunks] index of the chunk's first |d|^2 partial
    const int* pv_idx;
    const double* pv_xyz;
    const int* pv_start;
    long long pv_stride;
    int* nn_idx;                 // [total points]
    int* cnt;                    // [batch][V]
    unsigned long long* sum;     // [batch][V][3] fixed point 2^36
    double* qpart;               // [total q blocks]
    int* range_flag;             // [batch]
};

struct LmArgs {
    double* x;                   // [batch][nx] in/out
    const int* cnt;
    const unsigned long long* sum;
    const double* qpart;
    const int* frame_qblock;     // [batch+1]
    double* Hcur;                // [batch][P*P] scratch
    double beta_pose, beta_shape, function_tolerance;
    int max_iters;
    double* dump_cost;           // nullable [batch]
    double* dump_grad;           // [batch][P]
    double* dump_H;              // [batch][P*P]
    double* trace;               // nullable [batch][trace_cap][nx]
    int trace_cap;
    FrameStats* stats;           // [batch]
    const int* range_flag;       // [batch]
};

size_t pose_smem_bytes(int V, int J, int K);
cudaError_t launch_pose_visibility(const DevModel& M, const DevParts& Pt, const PoseArgs& a, int batch, cudaStream_t st);
cudaError_t launch_nn(const DevParts& Pt, const NNArgs& a, int num_chunks, cudaStream_t st);
cudaError_t launch_lm(const DevModel& M, const DevParts& Pt, const LmArgs& a, int batch, bool acc64, cudaStream_t st);

}  // namespace avb
