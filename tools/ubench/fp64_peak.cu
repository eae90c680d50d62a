// fp64_peak.cu -- measured fp64 peaks of this GPU, the denominators of bench.py's `roofline_fp64`:
//   dfma : register-resident chains of DFMA (8 independent accumulators per thread), 148 x 4 CTAs x 256 threads
//   dmma : register-resident mma.sync.m8n8k4.f64 chains (8 independent accumulator pairs per warp)
// Prints one JSON object (TFLOP/s = 2 x multiply-adds).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 4, iters = 20000;
    double* d;
    cudaMalloc(&d, (size_t)grid * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_f = 0, ms_m = 0;
    for (int rep = 0; rep < 3; ++rep) {
        k_dfma<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dfma<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0 || ms < ms_f) ms_f = ms;
        k_dmma<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dmma<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0 || ms < ms_m) ms_m = ms;
    }
    const double fma_ops = (double)grid * 256 * iters * 32.0;                 // DFMA per launch
    const double mma_ops = (double)grid * 8 * iters * 32.0 * (8 * 8 * 4);     // multiply-adds per launch
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, "
           "\"how\": \"register-resident chains, 4 CTAs x 256 threads per SM, best of 3, CUDA events; TFLOP/s = 2 x multiply-adds / s\", "
           "\"cuda\": \"%s\"}\n",
           prop.name, prop.multiProcessorCount, 2 * fma_ops / ms_f * 1e-9, 2 * mma_ops / ms_m * 1e-9, ms_f, ms_m,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
