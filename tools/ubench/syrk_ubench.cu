// Micro-benchmark of the A^T A inner loop: which (tile shape, smem element type, warps/SM) reaches the fp64 pipe peak?
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <typename SmemT, int TI, int TJ, int THREADS>
__global__ void __launch_bounds__(THREADS) k_syrk(const float* __restrict__ src, double* out, int rows, int lda, int iters, int nrg) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmemT* A = reinterpret_cast<SmemT*>(raw);
    for (int e = threadIdx.x; e < rows * lda; e += THREADS) A[e] = (SmemT)src[e % 4096];
    __syncthreads();
    const int n_i = (lda - 4) / TI, n_j = (lda - 4) / TJ;
    const int pair = threadIdx.x / nrg, rg = threadIdx.x % nrg;
    const int bi = pair % n_i, bj = (pair / n_i) % n_j;
    double acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int r = rg; r < rows; r += nrg) {
            double av[TI], bv[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) av[i] = (double)A[r * lda + TI * bi + i];
#pragma unroll
            for (int j = 0; j < TJ; ++j) bv[j] = (double)A[r * lda + TJ * bj + j];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) s += acc[i][j];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <typename SmemT, int TI, int TJ, int THREADS>
void run(const char* name, int ctas_per_sm, int nrg, const float* d_src, double* d_out) {
    const int rows = 192, lda = 60, iters = 200;
    const size_t smem = (size_t)rows * lda * sizeof(SmemT);
    cudaFuncSetAttribute(k_syrk<SmemT, TI, TJ, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_syrk<SmemT, TI, TJ, THREADS><<<grid, THREADS, smem>>>(d_src, d_out, rows, lda, 2, nrg);
    cudaEventRecord(e0);
    k_syrk<SmemT, TI, TJ, THREADS><<<grid, THREADS, smem>>>(d_src, d_out, rows, lda, iters, nrg);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)grid * THREADS * iters * (rows / (double)nrg) * TI * TJ;
    printf("%-34s ctas/sm=%d nrg=%2d  %.3f ms  %.2f TDFMA/s (%.1f%% of 18.6)  err=%s\n", name, ctas_per_sm, nrg, ms, dfma / ms * 1e-9,
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
    for (int nrg : {1, 4, 32}) {
        run<float, 8, 4, 256>("f32 smem, 8x4, 256 thr", 2, nrg, d_src, d_out);
        run<double, 8, 4, 256>("f64 smem, 8x4, 256 thr", 2, nrg, d_src, d_out);
        run<float, 8, 8, 128>("f32 smem, 8x8, 128 thr", 2, nrg, d_src, d_out);
        run<double, 8, 8, 128>("f64 smem, 8x8, 128 thr", 2, nrg, d_src, d_out);
        run<double, 8, 8, 256>("f64 smem, 8x8, 256 thr", 1, nrg, d_src, d_out);
        run<double, 4, 4, 256>("f64 smem, 4x4, 256 thr", 4, nrg, d_src, d_out);
        run<float, 4, 4, 256>("f32 smem, 4x4, 256 thr", 4, nrg, d_src, d_out);
    }
    return 0;
}
