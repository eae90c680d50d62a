// Micro-benchmark of the latency-bound pieces of lm_solve (blocked Cholesky, back substitution) in isolation: one CTA of
// 256 threads per SM factors a synthetic SPD 85x85 system, alone and next to a CTA that saturates the fp64 pipe with
// DMMA (what the co-resident Gram task of lm_flow_kernel does).  Cycle counts per phase from clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -o solve_ubench solve_ubench.cu
#include <cstdio>
#include <vector>
#define AVB_UBENCH_PHASES
#include "../../avatar_b200/csrc/avb_lm.cu"

using namespace avb;

// role test: CTAs < 148 time back substitutions while their SM neighbour (CTA + 148) factors continuously
__global__ void __launch_bounds__(256, 2) k_roles(const double* A, int P, long long* cyc, int reps, double* out) {
    extern __shared__ __align__(16) unsigned char raw[];
    double* W = reinterpret_cast<double*>(raw);
    double* dinv = W + (size_t)(P + 2) * P;
    double* wscr = dinv + 128;
    double* x = wscr + 8 * 64 + 8 * 96 + 8;
    for (int i = threadIdx.x; i < (P + 1) * P; i += 256) W[i] = A[i];
    __syncthreads();
    if (blockIdx.x >= 148) {
        for (int r = 0; r < reps; ++r) {
            for (int i = threadIdx.x; i < (P + 1) * P; i += 256) W[i] = A[i];
            __syncthreads();
            aug_cholesky(W, P, P + 1, dinv, wscr, wscr + 8 * 64);
        }
        return;
    }
    aug_cholesky(W, P, P + 1, dinv, wscr, wscr + 8 * 64);
    long long tb = 0;
    for (int r = 0; r < reps * 3; ++r) {
        long long t1 = clock64();
        if (threadIdx.x < 32) warp_back_solve(W, dinv, P, W + (size_t)P * P, x);
        __syncthreads();
        tb += clock64() - t1;
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = tb / (reps * 3);
    if (blockIdx.x == 0) out[threadIdx.x & 63] = x[threadIdx.x & 63];
}

__global__ void __launch_bounds__(256, 2) k_solve(const double* A, int P, long long* cyc, int reps, double* out) {
    extern __shared__ __align__(16) unsigned char raw[];
    double* W = reinterpret_cast<double*>(raw);   // (P+1) x P
    double* dinv = W + (size_t)(P + 2) * P;
    double* wscr = dinv + 128;
    double* x = wscr + 8 * 64 + 8 * 96 + 8;
    long long tc = 0, tb = 0, tc0 = 0, tb0 = 0;
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < (P + 1) * P; i += 256) W[i] = A[i];
        __syncthreads();
        long long t0 = clock64();
        bool ok = aug_cholesky(W, P, P + 1, dinv, wscr, wscr + 8 * 64);
        long long t1 = clock64();
        if (threadIdx.x < 32) warp_back_solve(W, dinv, P, W + (size_t)P * P, x);
        __syncthreads();
        long long t2 = clock64();
        if (r == 0) { tc0 = t1 - t0; tb0 = t2 - t1; } else { tc += t1 - t0; tb += t2 - t1; }
        if (!ok) break;
    }
    if (threadIdx.x == 0) {
        cyc[4 * blockIdx.x] = tc / (reps - 1);
        cyc[4 * blockIdx.x + 1] = tb / (reps - 1);
        cyc[4 * blockIdx.x + 2] = tc0;
        cyc[4 * blockIdx.x + 3] = tb0;
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < P; i += 256) out[i] = x[i];
}

__global__ void __launch_bounds__(256, 2) k_hog(double* out, int iters) {
    double acc[8][2];
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// hogs that imitate what a co-resident record / Gram task issues: kind 1 = DFMA stream, 2 = LDS.64 stream,
// 3 = LDS.32 + F2F + DMMA (the Gram inner loop), 4 = global loads (L2 hits)
__global__ void __launch_bounds__(256, 2) k_hog2(double* out, const double* src, int iters, int kind) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = 1.0 + i * 1e-6;
    __syncthreads();
    double acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    const float* smf = reinterpret_cast<const float*>(sm);
    for (int it = 0; it < iters; ++it) {
        if (kind == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], 1.0000001, 1e-9);
        } else if (kind == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += sm[(threadIdx.x * 3 + it * 7 + i * 33) & 2047];
        } else if (kind == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = (double)smf[(threadIdx.x + it * 4 + i * 68) & 4095], b = (double)smf[(threadIdx.x + 17 + it * 4 + i * 68) & 4095];
                dmma884(acc[2 * i], acc[2 * i + 1], a, b);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __ldcg(src + ((threadIdx.x + (it * 8 + i) * 256 + blockIdx.x * 4096) & 0xFFFFF));
        }
    }
    double s2 = 0;
    for (int i = 0; i < 8; ++i) s2 += acc[i];
    out[blockIdx.x * 256 + threadIdx.x] = s2;
}

int main() {
    const int P = 85;
    std::vector<double> A((size_t)(P + 1) * P), B((size_t)P * P);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
    for (auto& v : B) v = rnd();
    for (int i = 0; i < P; ++i)
        for (int j = 0; j < P; ++j) {
            double t = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < P; ++k) t += B[i * P + k] * B[j * P + k];
            A[i * P + j] = t;
        }
    for (int j = 0; j < P; ++j) A[(size_t)P * P + j] = rnd();
    double *dA, *dout, *dh;
    long long* dc;
    cudaMalloc(&dA, A.size() * 8);
    cudaMalloc(&dout, 1024 * 8);
    cudaMalloc(&dh, 148 * 2 * 256 * 8);
    cudaMalloc(&dc, 148 * 2 * 4 * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = ((size_t)(P + 2) * P + 128 + 512 + 8 * 96 + 8 + 128) * 8;
    cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t s1, s2;
    cudaStreamCreate(&s1);
    cudaStreamCreate(&s2);
    double* dsrc; cudaMalloc(&dsrc, (1 << 20) * 8); cudaMemset(dsrc, 0, (1 << 20) * 8);
    for (int mode = 0; mode < 7; ++mode) {
        // mode 0: one solve CTA per SM alone; 1: two solve CTAs per SM; 2: one solve CTA + one DMMA hog CTA per SM
        const int grid = mode == 1 ? 296 : 148;
        if (mode == 2) k_hog<<<148, 256, 0, s2>>>(dh, 400000);
        if (mode >= 3) k_hog2<<<148, 256, 0, s2>>>(dh, dsrc, mode == 6 ? 30000 : 600000, mode - 2);
        k_solve<<<grid, 256, smem, s1>>>(dA, P, dc, 20, dout);
        cudaStreamSynchronize(s1);
        std::vector<long long> c(4 * grid);
        cudaMemcpy(c.data(), dc, c.size() * 8, cudaMemcpyDeviceToHost);
        double mc = 0, mb = 0, mc0 = 0, mb0 = 0;
        for (int i = 0; i < grid; ++i) { mc += c[4 * i]; mb += c[4 * i + 1]; mc0 += c[4 * i + 2]; mb0 += c[4 * i + 3]; }
        printf("mode %d: cholesky %.0f cycles, back substitution %.0f cycles warm; first execution %.0f / %.0f (mean over %d CTAs)  err=%s\n",
               mode, mc / grid, mb / grid, mc0 / grid, mb0 / grid, grid, cudaGetErrorString(cudaGetLastError()));
        cudaDeviceSynchronize();
        long long pc[4];
        cudaMemcpyFromSymbol(pc, avb::g_chol_cyc, 32);
        printf("   CTA 0 per factorisation: diag %lld, rows %lld, trailing %lld cycles\n", pc[0] / 20, pc[1] / 20, pc[2] / 20);
        long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(avb::g_chol_cyc, z, 32);
    }
    {
        cudaFuncSetAttribute(k_roles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_roles<<<296, 256, smem, s1>>>(dA, P, dc, 20, dout);
        cudaStreamSynchronize(s1);
        std::vector<long long> c(148);
        cudaMemcpy(c.data(), dc, c.size() * 8, cudaMemcpyDeviceToHost);
        double mb = 0;
        for (int i = 0; i < 148; ++i) mb += c[i];
        printf("roles: back substitution %.0f cycles next to a CTA that factors continuously  err=%s\n", mb / 148, cudaGetErrorString(cudaGetLastError()));
        k_solve<<<148, 256, smem, s1>>>(dA, P, dc, 20, dout);
        cudaStreamSynchronize(s1);
    }
    std::vector<double> x(P);
    cudaMemcpy(x.data(), dout, P * 8, cudaMemcpyDeviceToHost);
    // residual check: A x = b
    double err = 0;
    for (int i = 0; i < P; ++i) {
        double t = 0;
        for (int j = 0; j < P; ++j) t += A[i * P + j] * x[j];
        err = fmax(err, fabs(t - A[(size_t)P * P + i]));
    }
    printf("max |A x - b| = %.3e\n", err);
    return 0;
}
