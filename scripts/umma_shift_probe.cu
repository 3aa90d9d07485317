// umma_shift_probe.cu -- does a K-major SWIZZLE_128B UMMA descriptor address a tile that starts at a row
// that is NOT a multiple of 8 (i.e. a start address that is 128-byte but not 1024-byte aligned)?
// This is what a sliding-window (halo) A operand for 3x3 convolutions needs: one TMA load of Wt+2 pixels,
// three MMAs whose A descriptors start at pixel 0, 1 and 2.
//
// Test: A_full = [136 rows][64 bf16] loaded by one 2-D TMA (SWIZZLE_128B) into 1024-aligned smem.
// For shift r in 0..7 and each base_offset candidate, D = A_full[r : r+128] * B^T (B = [128][64], K = 64)
// via tcgen05.mma with the A start address advanced by r*128 bytes; compare with a host reference.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe umma_shift_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *out,
                                             int shift, int base_off_mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                 // 136 x 128 B = 17408 -> pad to 18432
    uint8_t *sB = smem + 18432;         // 128 x 128 B = 16384
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 16384);
    uint64_t *mbar = bar + 1;
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tslot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(136 * 128 + 128 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s_addr(sA)),
                     "l"(&tmA), "r"(s_addr(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s_addr(sB)),
                     "l"(&tmB), "r"(s_addr(bar)), "r"(0), "r"(0) : "memory");
    }
    // everyone waits for the data
    {
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s_addr(bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = s_addr(sA) + (uint32_t)shift * 128u;
        uint64_t base_off = 0;
        if (base_off_mode == 1) base_off = (a_addr >> 7) & 7;
        const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t a_desc = hi | (uint64_t)((a_addr & 0x3FFFF) >> 4) | (base_off << 49);
        const uint64_t b_desc = hi | (uint64_t)((s_addr(sB) & 0x3FFFF) >> 4);
        for (int ks = 0; ks < 4; ++ks) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(a_desc + (uint64_t)(ks * 2)), "l"(b_desc + (uint64_t)(ks * 2)), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_addr(mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s_addr(mbar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread t reads TMEM lane t (row t), 128 columns
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                       "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 128 + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int RA = 136, RB = 128, K = 64;
    std::vector<__nv_bfloat16> hA(RA * K), hB(RB * K);
    std::vector<float> fA(RA * K), fB(RB * K);
    srand(1);
    for (int i = 0; i < RA * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < RB * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float *dO;
    CK(cudaMalloc(&dA, RA * K * 2)); CK(cudaMalloc(&dB, RB * K * 2)); CK(cudaMalloc(&dO, 128 * 128 * 4));
    CK(cudaMemcpy(dA, hA.data(), RA * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), RB * K * 2, cudaMemcpyHostToDevice));
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN_enc enc = (PFN_enc)fp;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)RA}; cuuint64_t str[1] = {(cuuint64_t)K * 2}; cuuint32_t box[2] = {64, (cuuint32_t)RA}; cuuint32_t es[2] = {1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
        cuuint64_t dimsB[2] = {(cuuint64_t)K, (cuuint64_t)RB}; cuuint32_t boxB[2] = {64, (cuuint32_t)RB};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
    }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    std::vector<float> hO(128 * 128);
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = 0; shift < 8; ++shift) {
            CK(cudaMemset(dO, 0, 128 * 128 * 4));
            probe<<<1, 128, 40 * 1024>>>(tmA, tmB, dO, shift, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hO.data(), dO, 128 * 128 * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 128; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)fA[(r + shift) * K + k] * fB[n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - hO[r * 128 + n]));
                }
            printf("base_offset_mode=%d shift=%d rows: max |err| = %.4g  %s\n", mode, shift, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
