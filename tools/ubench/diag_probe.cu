// One warp factors an 8 x 8 SPD block held one row per lane (the diagonal-block step of the blocked Cholesky): cycle counts of
// candidate formulations.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o diag_probe diag_probe.cu
#include <cstdio>
#include <cmath>
#include <cstdint>
constexpr int kNB = 8;
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y = (double)rsqrtf((float)d);
    const double h = 0.5 * d;
    y = fma(y, fma(-h * y, y, 0.5), y);
    y = fma(y, fma(-h * y, y, 0.5), y);
    return y;
}
__device__ __forceinline__ double rcp_nr3(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}
// 1/d: 2^-20 seed, one cubic step (error e^3 = 2^-60), three dependent fp64 ops
__device__ __forceinline__ double rcp_cubic(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}
// 1/sqrt(d): 2^-19 seed, one cubic step: y (1 + e/2 + 3 e^2 / 8), e = 1 - d y^2 (error ~ e^3 = 2^-57) + one fix-up
__device__ __forceinline__ double rsqrt_fast(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d * y, y, 1.0);
    y = fma(y, fma(0.375 * e, e, 0.5 * e), y);
    e = fma(-d * y, y, 1.0);
    return fma(0.5 * y, e, y);
}
template <int V>
__device__ __forceinline__ bool diag_factor(double* W, int j0, int nb, double* Lw, double* dinv, int lane, long long* tm) {
    const int r = lane & 7;
    double a[kNB];
    long long t0 = clock64();
#pragma unroll
    for (int c = 0; c < kNB; ++c) a[c] = (r < nb && c <= r) ? W[tri(j0 + r) + j0 + c] : ((c == r) ? 1.0 : 0.0);
    bool ok = true;
    double mydiag = 1.0, myrs = 1.0;
    long long t1 = clock64();
    if (V == 0) {
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            const double d = __shfl_sync(0xffffffffu, a[k], k);
            if (!(d > 1e-300) || !(d < 1e300)) ok = false;
            const double inv = rcp_nr3(d);
            const double t = a[k] * inv;
            if (r == k) mydiag = d;
#pragma unroll
            for (int c = k + 1; c < kNB; ++c) {
                const double ack = __shfl_sync(0xffffffffu, a[k], c);
                a[c] = fma(-t, ack, a[c]);
            }
            a[k] = t;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            const double d = __shfl_sync(0xffffffffu, a[k], k);
            if (!(d > 1e-300) || !(d < 1e300)) ok = false;
            const double inv = rcp_cubic(d);
#pragma unroll
            for (int c = k + 1; c < kNB; ++c) {
                const double ack = __shfl_sync(0xffffffffu, a[k], c);
                a[c] = fma(-(a[k] * ack), inv, a[c]);
            }
            const double rs = rsqrt_fast(d);   // off the pivot chain
            if (r == k) { mydiag = d; myrs = rs; }
            a[k] = a[k] * rs;                  // L_rk = a_rk / sqrt(d_k)   (lane k: sqrt(d_k))
        }
    }
    long long t2 = clock64();
    if (!ok) return false;
    if (V == 0) {
        const double rs = pivot_rsqrt(mydiag), sq = mydiag * rs;
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            const double sqk = __shfl_sync(0xffffffffu, sq, k);
            const double l = (k < r) ? a[k] * sqk : ((k == r) ? sq : 0.0);
            if (lane < kNB) {
                Lw[r * kNB + k] = (k < r) ? l * rs : ((k == r) ? rs : 0.0);
                if (r < nb && k <= r) W[tri(j0 + r) + j0 + k] = l;
            }
        }
        if (lane < nb) dinv[j0 + lane] = rs;
    } else {
        if (lane < kNB) {
#pragma unroll
            for (int k = 0; k < kNB; ++k) {
                Lw[r * kNB + k] = (k < r) ? a[k] * myrs : ((k == r) ? myrs : 0.0);
                if (r < nb && k <= r) W[tri(j0 + r) + j0 + k] = a[k];
            }
        }
        if (lane < nb) dinv[j0 + lane] = myrs;
    }
    __syncwarp();
    long long t3 = clock64();
    if (lane == 0) { tm[0] = t1 - t0; tm[1] = t2 - t1; tm[2] = t3 - t2; }
    return true;
}
template <int V>
__global__ void k(const double* A, double* out, long long* tm) {
    __shared__ double W[64], Lw[64], dinv[8];
    for (int rep = 0; rep < 3; ++rep) {
        for (int i = threadIdx.x; i < 36; i += 32) W[i] = A[i];
        __syncwarp();
        diag_factor<V>(W, 0, 8, Lw, dinv, threadIdx.x, tm);
        __syncwarp();
    }
    for (int i = threadIdx.x; i < 36; i += 32) out[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += 32) out[36 + i] = Lw[i];
    if (threadIdx.x < 8) out[100 + threadIdx.x] = dinv[threadIdx.x];
}
int main() {
    double A[36], B[64];
    unsigned s = 99;
    for (auto& v : B) { s = s * 1664525u + 1013904223u; v = (double)(s >> 8) / (1 << 24) - 0.5; }
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j <= i; ++j) {
            double t = (i == j) ? 0.5 : 0.0;
            for (int k = 0; k < 8; ++k) t += B[i * 8 + k] * B[j * 8 + k];
            A[i * (i + 1) / 2 + j] = t;
        }
    double L[36];   // host Cholesky
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j <= i; ++j) {
            long double t = A[i * (i + 1) / 2 + j];
            for (int k = 0; k < j; ++k) t -= (long double)L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            L[i * (i + 1) / 2 + j] = (i == j) ? (double)sqrtl(t) : (double)(t / L[j * (j + 1) / 2 + j]);
        }
    double *dA, *dout; long long* dt;
    cudaMalloc(&dA, 36 * 8); cudaMalloc(&dout, 128 * 8); cudaMalloc(&dt, 64);
    cudaMemcpy(dA, A, 36 * 8, cudaMemcpyHostToDevice);
    for (int v = 0; v < 2; ++v) {
        if (v == 0) k<0><<<1, 32>>>(dA, dout, dt); else k<1><<<1, 32>>>(dA, dout, dt);
        double o[128]; long long tm[4];
        cudaMemcpy(o, dout, 128 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(tm, dt, 24, cudaMemcpyDeviceToHost);
        double worst = 0, wl = 0;
        for (int i = 0; i < 36; ++i) worst = fmax(worst, fabs(o[i] - L[i]) / fabs(L[i]));
        for (int i = 0; i < 8; ++i) {
            wl = fmax(wl, fabs(o[100 + i] * L[i * (i + 1) / 2 + i] - 1.0));
            for (int j = 0; j <= i; ++j) {
                const double want = (j == i) ? 1.0 / L[i * (i + 1) / 2 + i] : L[i * (i + 1) / 2 + j] / L[i * (i + 1) / 2 + i];
                wl = fmax(wl, fabs(o[36 + i * 8 + j] - want) / fabs(want));
            }
        }
        printf("variant %d: load %lld, pivots %lld, tail %lld cycles; max rel err L %.2e, Lw/dinv %.2e  %s\n", v, tm[0], tm[1], tm[2], worst, wl, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
