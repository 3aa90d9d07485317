// umma_2cta_probe.cu -- minimal cta_group::2 (CTA pair) tcgen05 GEMM: D[256 x 128] = A[256 x 64] * B[128 x 64]^T, bf16 -> fp32.
// A cluster of two CTAs; CTA r loads A rows [128r, 128r+128) and B rows [64r, 64r+64) with 2-SM TMA loads that signal the
// LEADER's mbarrier; the leader issues tcgen05.mma.cta_group::2 (M = 256) and commits with a multicast arrive to both
// CTAs; each CTA reads its own 128 accumulator rows from its TMEM.  Checks the result against the host.
// Purpose: pin down the PTX details (allocation, remote barrier signalling, multicast commit) before the conv kernel uses them.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_2cta_probe umma_2cta_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                 // 128 x 128 B = 16 KB
    uint8_t *sB = smem + 16384;         // 64 x 128 B = 8 KB
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + 8192);
    uint64_t *done = full + 1;
    uint32_t *tslot = reinterpret_cast<uint32_t *>(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tslot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();                      // both CTAs' barriers exist before anybody signals a remote one
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (threadIdx.x == 0) {
        // leader's full barrier, addressed from either CTA
        uint32_t lead_full;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lead_full) : "r"(s_addr(full)), "r"(0));
        if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(full)), "r"(2 * (16384 + 8192)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s_addr(sA)),
                     "l"(&tmA), "r"(lead_full), "r"(0), "r"((int)rank * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s_addr(sB)),
                     "l"(&tmB), "r"(lead_full), "r"(0), "r"((int)rank * 64) : "memory");
    }
    if (rank == 0 && warp == 1) {
        mbar_wait(s_addr(full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
            const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const uint64_t a_desc = hi | (uint64_t)((s_addr(sA) & 0x3FFFF) >> 4);
            const uint64_t b_desc = hi | (uint64_t)((s_addr(sB) & 0x3FFFF) >> 4);
            for (int ks = 0; ks < 4; ++ks)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                             "l"(a_desc + (uint64_t)(ks * 2)), "l"(b_desc + (uint64_t)(ks * 2)), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s_addr(done)),
                         "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(s_addr(done), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                       "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((int)rank * 128 + warp * 32 + lane) * 128 + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
    if (threadIdx.x == 0 && blockIdx.x < 2) out[256 * 128 + rank] = (float)tmem;     // report the TMEM base of each CTA
}

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int M = 256, N = 128, K = 64;
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(2);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float *dO;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dO, (M * N + 8) * 4));
    CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0, (M * N + 8) * 4));
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN_enc enc = (PFN_enc)fp;
    CUtensorMap tmA, tmB;
    cuuint64_t str[1] = {(cuuint64_t)K * 2}; cuuint32_t es[2] = {1, 1};
    cuuint64_t dimsA[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint32_t boxA[2] = {64, 128};
    cuuint64_t dimsB[2] = {(cuuint64_t)K, (cuuint64_t)N}; cuuint32_t boxB[2] = {64, 64};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    probe2<<<2, 128, 32 * 1024>>>(tmA, tmB, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> hO(M * N + 8);
    CK(cudaMemcpy(hO.data(), dO, (M * N + 8) * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad_row = -1;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[r * K + k] * fB[n * K + k];
            const double e = fabs(ref - hO[r * N + n]);
            if (e > maxerr) { maxerr = e; bad_row = r; }
        }
    printf("cta_group::2 M=256 N=128 K=64: max |err| = %.4g (row %d)  %s;  tmem base cta0=%g cta1=%g\n", maxerr, bad_row, maxerr < 1e-3 ? "OK" : "MISMATCH",
           hO[M * N], hO[M * N + 1]);
    return 0;
}
