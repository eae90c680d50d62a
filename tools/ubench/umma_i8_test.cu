// umma_i8_test.cu -- stand-alone check of the EXACT integer Gram matrix on the tcgen05 int8 path (B200 / sm_100a):
// every field value is a 22-bit fixed-point integer q = t0 * 2^14 + t1 * 2^7 + t2 with int8 slices |t0| <= 128 (clamped to
// 127), |t1|, |t2| <= 64; slice tiles [80 rows][128 vertices] of int8 in the K-major SWIZZLE_128B layout (one swizzle atom:
// 128 bytes = 128 vertices per row); tcgen05.mma.kind::i8 (M = 128, N = 80, K = 32) accumulates the slice products of equal
// weight class w = a + b into three int32 TMEM accumulators (columns 0, 80, 160).  Integer accumulation is exact, so
// G = A0 * 2^28 + A1 * 2^21 + A2 * 2^14 (+ classes 3, 4 dropped) must equal the CPU integer result bit for bit.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int kRows = 80, kKv = 128, kSliceBytes = kRows * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int FMT>   // a/b format code of the instruction descriptor under test
__global__ void __launch_bounds__(256) gram_i8(const int* __restrict__ Q, int* __restrict__ A, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int k = (tid & 15) + 16 * (tid >> 5), h = (tid >> 4) & 1;
    for (int m = h; m < kRows; m += 2) {
        const int q = Q[m * kKv + k];
        int t0 = (q + (1 << 13)) >> 14;
        int rem = q - (t0 << 14);
        int t1 = (rem + (1 << 6)) >> 7;
        int t2 = rem - (t1 << 7);
        const uint32_t off = (uint32_t)(m * 128 + ((((k >> 4) ^ m) & 7) << 4) + (k & 15));
        tile[off] = (unsigned char)(signed char)t0;
        tile[kSliceBytes + off] = (unsigned char)(signed char)t1;
        tile[2 * kSliceBytes + off] = (unsigned char)(signed char)t2;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_d = tmem_base_s;
    long long t0c = clock64();
    if (tid == 0) {
        // D = S32 (2 at [4,6)), A = B = FMT at [7,10) / [10,13), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
        const uint32_t idesc = (2u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t base = smem_u32(tile);
        for (int w = 0; w < 3; ++w) {
            uint32_t accum = 0;
            for (int ta = 0; ta <= w; ++ta) {
                const int tb = w - ta;
                for (int ks = 0; ks < kKv / 32; ++ks) {
                    const uint64_t da = make_desc(base + ta * kSliceBytes + ks * 32), db = make_desc(base + tb * kSliceBytes + ks * 32);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem_d + (uint32_t)(80 * w)), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                    accum = 1;
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    long long t1c = clock64();
    if (tid == 0 && cycles) *cycles = t1c - t0c;
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int m = 32 * (wid & 3) + lane;
    for (int cb = (wid >> 2) * 120; cb < (wid >> 2) * 120 + 120; cb += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)cb;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < kRows)
            for (int j = 0; j < 8; ++j) A[m * 240 + cb + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256));
}

template <int FMT>
void run(const std::vector<int>& hQ, int* dQ, int* dA, long long* dC) {
    const size_t smem = 3 * kSliceBytes + 1024 + 8192;
    cudaFuncSetAttribute(gram_i8<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(dA, 0, kRows * 240 * 4);
    gram_i8<FMT><<<1, 256, smem>>>(dQ, dA, dC);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> a(kRows * 240);
    long long cyc = 0;
    cudaMemcpy(a.data(), dA, a.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    // CPU: the same slices, classes 0..2
    long long bad = 0;
    double worst_rel = 0;
    std::vector<int> t(3 * kRows * kKv);
    for (int m = 0; m < kRows; ++m)
        for (int k = 0; k < kKv; ++k) {
            const int q = hQ[m * kKv + k];
            int t0 = (q + (1 << 13)) >> 14, rem = q - (t0 << 14), t1 = (rem + (1 << 6)) >> 7, t2 = rem - (t1 << 7);
            t[(0 * kRows + m) * kKv + k] = t0; t[(1 * kRows + m) * kKv + k] = t1; t[(2 * kRows + m) * kKv + k] = t2;
        }
    for (int i = 0; i < kRows; ++i)
        for (int j = 0; j < kRows; ++j) {
            long long cls[3] = {0, 0, 0}, exact = 0;
            for (int k = 0; k < kKv; ++k) {
                exact += (long long)hQ[i * kKv + k] * hQ[j * kKv + k];
                for (int ta = 0; ta < 3; ++ta)
                    for (int tb = 0; ta + tb < 3; ++tb) cls[ta + tb] += (long long)t[(ta * kRows + i) * kKv + k] * t[(tb * kRows + j) * kKv + k];
            }
            for (int w = 0; w < 3; ++w) bad += (cls[w] != a[i * 240 + 80 * w + j]);
            const long long g = ((long long)a[i * 240 + j] << 28) + ((long long)a[i * 240 + 80 + j] << 21) + ((long long)a[i * 240 + 160 + j] << 14);
            double dii = 0, djj = 0;
            for (int k = 0; k < kKv; ++k) { dii += (double)hQ[i * kKv + k] * hQ[i * kKv + k]; djj += (double)hQ[j * kKv + k] * hQ[j * kKv + k]; }
            worst_rel = std::fmax(worst_rel, std::fabs((double)(g - exact)) / (std::sqrt(dii * djj) + 1e-300));
        }
    printf("{\"kind\": \"i8\", \"fmt_code\": %d, \"class_mismatches\": %lld, \"max_err_vs_exact_over_diag_scale\": %.3e, \"mma_cycles\": %lld, \"cuda\": \"%s\"}\n",
           FMT, bad, worst_rel, cyc, cudaGetErrorString(e));
}

int main() {
    std::vector<int> hQ(kRows * kKv);
    uint32_t s = 777u;
    for (auto& x : hQ) {
        s = s * 1664525u + 1013904223u;
        x = (int)(s >> 10) - (1 << 21);          // (-2^21, 2^21)
        if (x >= (127 << 14) + (1 << 13)) x = (127 << 14);   // keep t0 <= 127
    }
    int *dQ, *dA;
    long long* dC;
    cudaMalloc(&dQ, hQ.size() * 4);
    cudaMalloc(&dA, kRows * 240 * 4);
    cudaMalloc(&dC, 8);
    cudaMemcpy(dQ, hQ.data(), hQ.size() * 4, cudaMemcpyHostToDevice);
    run<1>(hQ, dQ, dA, dC);
    run<0>(hQ, dQ, dA, dC);
    return 0;
}
