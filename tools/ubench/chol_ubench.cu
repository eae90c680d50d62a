// Micro-benchmark of the packed-triangle blocked Cholesky of lm_solve (aug_cholesky, warp_back_solve) in isolation:
// one CTA of 256 threads per SM factors a synthetic SPD 85 x 85 system with one right-hand-side row; cycle counts per
// phase from clock64(), result checked against a host factorisation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -o chol_ubench chol_ubench.cu
#include <cstdio>
#include <cmath>
#include <vector>
#define AVB_UBENCH_PHASES
#include "../../avatar_b200/csrc/avb_lm.cu"

using namespace avb;
__global__ void __launch_bounds__(256, 2) k_solve(const double* A, int P, long long* cyc, int reps, double* out) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int nTri = tri(P);
    double* W = reinterpret_cast<double*>(raw);   // packed lower + row P
    double* dinv = W + ((nTri + P + 2 + 1) & ~1);
    double* wscr = dinv + 128;
    double* panel = wscr + 8 * 64;
    double* x = panel + 8 * 100 + 16;
    long long tc = 0, tb = 0, tc0 = 0, tb0 = 0;
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < nTri + P; i += 256) W[i] = A[i];
        __syncthreads();
        long long t0 = clock64();
        bool ok = aug_cholesky(W, P, P + 1, dinv, wscr, panel);
        long long t1 = clock64();
        block_inverses(W, dinv, P, panel);
        __syncthreads();
        if (threadIdx.x < 32) warp_back_solve(W, panel, P, W + nTri, x);
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

int main() {
    const int P = 85, nTri = P * (P + 1) / 2;
    std::vector<double> A((size_t)nTri + P), B((size_t)P * P), F((size_t)P * P);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
    for (auto& v : B) v = rnd();
    for (int i = 0; i < P; ++i)
        for (int j = 0; j <= i; ++j) {
            double t = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < P; ++k) t += B[i * P + k] * B[j * P + k];
            A[i * (i + 1) / 2 + j] = t;
            F[i * P + j] = F[j * P + i] = t;
        }
    std::vector<double> b(P);
    for (int j = 0; j < P; ++j) A[(size_t)nTri + j] = b[j] = rnd();
    double *dA, *dout;
    long long* dc;
    cudaMalloc(&dA, A.size() * 8);
    cudaMalloc(&dout, 1024 * 8);
    cudaMalloc(&dc, 148 * 2 * 4 * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = ((size_t)nTri + P + 4 + 128 + 512 + 8 * 100 + 16 + 128) * 8;
    cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode) {
        const int grid = (mode & 1) ? 296 : 148;
        k_solve<<<grid, 256, smem>>>(dA, P, dc, 20, dout);
        cudaDeviceSynchronize();
        std::vector<long long> c(4 * grid);
        cudaMemcpy(c.data(), dc, c.size() * 8, cudaMemcpyDeviceToHost);
        double mc = 0, mb = 0, mc0 = 0, mb0 = 0;
        for (int i = 0; i < grid; ++i) { mc += c[4 * i]; mb += c[4 * i + 1]; mc0 += c[4 * i + 2]; mb0 += c[4 * i + 3]; }
        printf("mode %d: cholesky %.0f cycles, back substitution %.0f cycles warm; first execution %.0f / %.0f (mean over %d CTAs)  err=%s\n",
               mode, mc / grid, mb / grid, mc0 / grid, mb0 / grid, grid, cudaGetErrorString(cudaGetLastError()));
        long long pc[4];
        cudaMemcpyFromSymbol(pc, avb::g_chol_cyc, 32);
        printf("   CTA 0 per factorisation: first diagonal block %lld, panels %lld cycles\n", pc[0] / 20, pc[1] / 20);
        long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(avb::g_chol_cyc, z, 32);
    }
    std::vector<double> x(P);
    cudaMemcpy(x.data(), dout, P * 8, cudaMemcpyDeviceToHost);
    double worst = 0;   // residual check: F x = b
    for (int i = 0; i < P; ++i) {
        double t = -b[i];
        for (int k = 0; k < P; ++k) t += F[i * P + k] * x[k];
        worst = fmax(worst, fabs(t));
    }
    printf("max |A x - b| = %.3e\n", worst);
    return 0;
}
