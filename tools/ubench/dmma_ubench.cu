// Micro-benchmark: Gram matrix R^T R of fp32 records with fp64 accumulation, DMMA (mma.sync.m8n8k4.f64) vs the SIMT
// 8x8 register block.  The tile is field-major T[q][t] (q = record field, t = vertex), ldt = 68 floats (conflict-free
// fragment loads: bank = 4 * (lane >> 2) + (lane & 3)).  A warp owns a TM x TN grid of 8x8 output blocks.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename SmemT, int TM, int TN, int THREADS>
__global__ void __launch_bounds__(THREADS) k_dmma(const float* __restrict__ src, double* out, int nblk, int iters) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmemT* T = reinterpret_cast<SmemT*>(raw);
    constexpr int ldt = 68;
    for (int e = threadIdx.x; e < nblk * 8 * ldt; e += THREADS) T[e] = (SmemT)src[e % 4096];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int bi0 = (wid * TM) % nblk, bj0 = (wid * TN + 1) % nblk;
    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int k0 = 0; k0 < 64; k0 += 4) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = (double)T[(((bi0 + i) % nblk) * 8 + gid) * ldt + k0 + tig];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = (double)T[(((bj0 + j) % nblk) * 8 + gid) * ldt + k0 + tig];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <typename SmemT, int TM, int TN, int THREADS>
void run(const char* name, int ctas_per_sm, float* d_src, double* d_out) {
    const int nblk = 12, iters = 400;
    const size_t smem = (size_t)nblk * 8 * 68 * sizeof(SmemT);
    cudaFuncSetAttribute(k_dmma<SmemT, TM, TN, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dmma<SmemT, TM, TN, THREADS><<<grid, THREADS, smem>>>(d_src, d_out, nblk, 2);
    cudaEventRecord(e0);
    k_dmma<SmemT, TM, TN, THREADS><<<grid, THREADS, smem>>>(d_src, d_out, nblk, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)grid * (THREADS / 32) * iters * 16.0 * TM * TN * 256.0;
    printf("%-40s ctas/sm=%d  %.3f ms  %.2f TDFMA/s (%.1f%% of SIMT peak 18.6)  err=%s\n", name, ctas_per_sm, ms, dfma / ms * 1e-9,
           dfma / ms * 1e-9 / 18.6 * 100, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* d_src;
    double* d_out;
    std::vector<float> h(4096);
    for (int i = 0; i < 4096; ++i) h[i] = (float)(i % 97) * 0.01f;
    cudaMalloc(&d_src, 4096 * 4);
    cudaMemcpy(d_src, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, 148 * 8 * 1024 * 8);
    run<float, 4, 4, 128>("dmma f32 tile, warp 4x4 blocks, 128 thr", 1, d_src, d_out);
    run<float, 4, 4, 128>("dmma f32 tile, warp 4x4 blocks, 128 thr", 2, d_src, d_out);
    run<float, 4, 4, 128>("dmma f32 tile, warp 4x4 blocks, 128 thr", 4, d_src, d_out);
    run<float, 4, 4, 256>("dmma f32 tile, warp 4x4 blocks, 256 thr", 2, d_src, d_out);
    run<double, 4, 4, 128>("dmma f64 tile, warp 4x4 blocks, 128 thr", 2, d_src, d_out);
    run<double, 4, 4, 128>("dmma f64 tile, warp 4x4 blocks, 128 thr", 4, d_src, d_out);
    run<float, 2, 4, 128>("dmma f32 tile, warp 2x4 blocks, 128 thr", 4, d_src, d_out);
    run<float, 3, 3, 128>("dmma f32 tile, warp 3x3 blocks, 128 thr", 4, d_src, d_out);
    run<float, 2, 2, 256>("dmma f32 tile, warp 2x2 blocks, 256 thr", 4, d_src, d_out);
    run<float, 4, 8, 128>("dmma f32 tile, warp 4x8 blocks, 128 thr", 2, d_src, d_out);
    run<double, 4, 8, 128>("dmma f64 tile, warp 4x8 blocks, 128 thr", 2, d_src, d_out);
    return 0;
}
