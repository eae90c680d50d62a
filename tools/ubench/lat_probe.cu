// Dependent-issue latencies (cycles per op in a serial chain, one warp) of the fp64 ops on the pivot chain of the
// blocked Cholesky, and the accuracy of the fp64 MUFU seeds.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cstdio>
#include <cmath>
__global__ void k(double* out, long long* cyc, double seed) {
    const int N = 512;
    double x = seed + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, 1e-9);
    long long t1 = clock64();
    double y = x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y));
    long long t2 = clock64();
    double z = y + x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
    long long t3 = clock64();
    double w = z;
#pragma unroll 16
    for (int i = 0; i < N; ++i) w = (double)rsqrtf((float)w) + 1.0;
    long long t4 = clock64();
    double v = w;
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(v));
    long long t5 = clock64();
    double u = v + 1.5;
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = u * 1.0000001;
    long long t6 = clock64();
    double s = u;
#pragma unroll 16
    for (int i = 0; i < N; ++i) s = s + 1e-9;
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = (t1 - t0) / N; cyc[1] = (t2 - t1) / N; cyc[2] = (t3 - t2) / N; cyc[3] = (t4 - t3) / N; cyc[4] = (t5 - t4) / N;
        cyc[5] = (t6 - t5) / N; cyc[6] = (t7 - t6) / N;
    }
    out[threadIdx.x] = x + y + z + w + v + u + s;
}
__global__ void kerr(double* err) {
    double worst_r = 0, worst_q = 0;
    unsigned s = 1234567u + threadIdx.x * 7919u;
    for (int i = 0; i < 20000; ++i) {
        s = s * 1664525u + 1013904223u;
        const double m = 1.0 + (double)(s >> 8) / (double)(1 << 24);
        s = s * 1664525u + 1013904223u;
        const double d = ldexp(m, (int)(s >> 26) - 32);
        double y, q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(d));
        worst_r = fmax(worst_r, fabs(fma(-d, y, 1.0)));
        worst_q = fmax(worst_q, fabs(fma(-d * q, q, 1.0)));
    }
    err[2 * threadIdx.x] = worst_r;
    err[2 * threadIdx.x + 1] = worst_q;
}
int main() {
    double* out; long long* cyc; double* err;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64); cudaMalloc(&err, 64 * 8);
    for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1.0);
    long long c[8];
    cudaMemcpy(c, cyc, 56, cudaMemcpyDeviceToHost);
    printf("cycles per dependent op: DFMA %lld, MUFU.RCP64H %lld, SHFL.64 %lld, F2F+RSQ+F2F+DADD %lld, MUFU.RSQ64H %lld, DMUL %lld, DADD %lld\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
    kerr<<<1, 32>>>(err);
    double e[64];
    cudaMemcpy(e, err, 64 * 8, cudaMemcpyDeviceToHost);
    double wr = 0, wq = 0;
    for (int i = 0; i < 32; ++i) { wr = fmax(wr, e[2 * i]); wq = fmax(wq, e[2 * i + 1]); }
    printf("seed error: |1 - d*rcp| <= %.3e (2^%.1f), |1 - d*rsq^2| <= %.3e (2^%.1f)  %s\n", wr, log2(wr), wq, log2(wq), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
