// umma_gram_test.cu -- stand-alone check of the tcgen05 building block of the fused record + Gram task of lm_flow_kernel:
// Gram matrix G = T T^T of a field-major tile T[nf][kv] (nf <= 80 record fields, kv = 128 vertices per sub-tile) with the
// fp32 fields split into NS bf16 terms (x = t0 + t1 + t2), every term tile written by ordinary shared-memory stores in
// the canonical K-major SWIZZLE_128B layout, products of total weight >= 2^-8(NS-1) issued as tcgen05.mma kind::f16
// (M = 128, N = 80, K = 16) into ONE fp32 TMEM accumulator, read back with tcgen05.ld.  Prints the error against an
// fp64 CPU Gram for NS = 1, 2, 3 and the cycles of one sub-tile's MMAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_gram_test umma_gram_test.cu && ./umma_gram_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int kRows = 80;            // record fields (padded to a multiple of 16 = UMMA N)
constexpr int kKv = 128;             // vertices per sub-tile = two 64-element swizzle atoms along K
constexpr int kAtomBytes = kRows * 128;      // one K-atom: kRows rows x 128 B (64 bf16)
constexpr int kTermBytes = 2 * kAtomBytes;   // one term tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row m, vertex k) inside one term tile, SWIZZLE_128B K-major: 8-row groups of 1024 B, the 16-byte
// chunk index of a row XORed with (row & 7)
__device__ __forceinline__ uint32_t tile_off(int m, int k) {
    const int atom = k >> 6, kk = k & 63;
    return (uint32_t)(atom * kAtomBytes + (m >> 3) * 1024 + (m & 7) * 128 + ((((kk >> 3) ^ (m & 7)) & 7) << 4) + ((kk & 7) << 1));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // start >> 4 | LBO (ignored for swizzled K-major; 1) | SBO = 1024 B (8-row group stride) | version 1 | SWIZZLE_128B (2 at [61,64))
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int NS>
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ T, float* __restrict__ G, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // two threads per vertex: half h writes the fields m = h, h + 2, ... (as the record task will)
    const int k = (tid & 15) + 16 * (tid >> 5), h = (tid >> 4) & 1;
    for (int m = h; m < kRows; m += 2) {
        float x = T[m * kKv + k];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const __nv_bfloat16 b = __float2bfloat16_rn(x);
            x -= __bfloat162float(b);
            *reinterpret_cast<__nv_bfloat16*>(tile + t * kTermBytes + tile_off(m, k)) = b;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_d = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t accum = 0;
#pragma unroll 1
        for (int ta = 0; ta < NS; ++ta)
#pragma unroll 1
            for (int tb = 0; tb + ta < NS; ++tb)   // products of weight 2^-8(ta+tb)
#pragma unroll 1
                for (int ks = 0; ks < kKv / 16; ++ks) {
                    const uint32_t koff = (uint32_t)((ks >> 2) * kAtomBytes + (ks & 3) * 32);
                    const uint64_t da = make_desc(smem_u32(tile + ta * kTermBytes) + koff);
                    const uint64_t db = make_desc(smem_u32(tile + tb * kTermBytes) + koff);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                    accum = 1;
                }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    long long t1 = clock64();
    if (tid == 0 && cycles) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::after_thread_sync;");
    // TMEM lane = row m; warp w reads lanes 32 (w & 3) .. +31, warps 0-3 the columns [0, 40), warps 4-7 [40, 80)
    const int m = 32 * (wid & 3) + lane, c0 = (wid >> 2) * 40;
    for (int cb = 0; cb < 40; cb += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)(c0 + cb);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < kRows)
            for (int j = 0; j < 8; ++j) G[m * kRows + c0 + cb + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128));
}

template <int NS>
void run(const std::vector<float>& hT, const std::vector<double>& ref, float* dT, float* dG, long long* dC) {
    const size_t smem = (size_t)NS * kTermBytes + 1024 + 8192;   // + slack: M = 128 reads 48 rows past the 80 of the last atom
    cudaFuncSetAttribute(gram_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(dG, 0, kRows * kRows * 4);
    gram_kernel<NS><<<1, 256, smem>>>(dT, dG, dC);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> g(kRows * kRows);
    long long cyc = 0;
    cudaMemcpy(g.data(), dG, g.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double worst = 0, asym = 0;
    for (int i = 0; i < kRows; ++i)
        for (int j = 0; j < kRows; ++j) {
            const double sc = std::sqrt(ref[i * kRows + i] * ref[j * kRows + j]) + 1e-300;
            worst = std::fmax(worst, std::fabs(g[i * kRows + j] - ref[i * kRows + j]) / sc);
            asym = std::fmax(asym, std::fabs((double)g[i * kRows + j] - g[j * kRows + i]) / sc);
        }
    printf("{\"terms\": %d, \"passes\": %d, \"max_rel_err\": %.3e, \"asymmetry\": %.3e, \"mma_cycles\": %lld, \"cuda\": \"%s\"}\n", NS,
           NS * (NS + 1) / 2, worst, asym, cyc, cudaGetErrorString(e));
}

int main() {
    std::vector<float> hT(kRows * kKv);
    uint32_t s = 12345u;
    for (auto& x : hT) {
        s = s * 1664525u + 1013904223u;
        x = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 4.0f;
    }
    for (int k = 0; k < kKv; ++k) hT[73 * kKv + k] = 0.f;   // rows >= nf are zero in the real tile
    std::vector<double> ref(kRows * kRows);
    for (int i = 0; i < kRows; ++i)
        for (int j = 0; j < kRows; ++j) {
            double a = 0;
            for (int k = 0; k < kKv; ++k) a += (double)hT[i * kKv + k] * hT[j * kKv + k];
            ref[i * kRows + j] = a;
        }
    float *dT, *dG;
    long long* dC;
    cudaMalloc(&dT, hT.size() * 4);
    cudaMalloc(&dG, kRows * kRows * 4);
    cudaMalloc(&dC, 8);
    cudaMemcpy(dT, hT.data(), hT.size() * 4, cudaMemcpyHostToDevice);
    run<1>(hT, ref, dT, dG, dC);
    run<2>(hT, ref, dT, dG, dC);
    run<3>(hT, ref, dT, dG, dC);
    return 0;
}
