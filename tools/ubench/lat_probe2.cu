// DMMA m8n8k4 latency / issue interval, MEMBAR.CTA and bar.sync cost.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe2 lat_probe2.cu
#include <cstdio>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc) {
    __shared__ double sm[256];
    const int N = 256;
    double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    sm[threadIdx.x] = a;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) dmma884(c0, c1, a, b);
    long long t1 = clock64();
    double d[8][2];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = 0;
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i)
#pragma unroll
        for (int u = 0; u < 8; ++u) dmma884(d[u][0], d[u][1], a, b);
    long long t2 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { sm[(threadIdx.x + i) & 255] = c0; __threadfence_block(); }
    long long t3 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t4 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { c0 += sm[(threadIdx.x + i) & 255]; }
    long long t5 = clock64();
    if (threadIdx.x == 0) { cyc[0] = (t1 - t0) / N; cyc[1] = (t2 - t1) / N; cyc[2] = (t3 - t2) / N; cyc[3] = (t4 - t3) / N; cyc[4] = (t5 - t4) / N; }
    double s = c0 + c1;
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
    out[threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 64);
    for (int nt = 32; nt <= 256; nt *= 8) {
        for (int r = 0; r < 2; ++r) k<<<1, nt>>>(out, cyc);
        long long c[8];
        cudaMemcpy(c, cyc, 40, cudaMemcpyDeviceToHost);
        printf("%d threads: DMMA dependent %lld cycles, 8 independent chains %lld per DMMA, STS+MEMBAR.CTA %lld, bar.sync %lld, LDS+DADD dependent %lld  %s\n", nt, c[0], c[1], c[2], c[3], c[4], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
