// What one SM's TMA unit delivers for the operand shapes of consistency_mma_kernel (B200).
//
// One CTA per SM, a ring of STAGES x 64 KB in shared memory; per stage the producer thread issues
// either   (mode 0) the kernel's six 2-D tensor boxes: rows of 128 bytes out of a plane with a
//                   20 096-byte pitch (2 x 128 rows + 4 x 64 rows, SWIZZLE_128B), or
//          (mode 1) the same 64 KB as 1-D bulk copies of CONTIGUOUS blocks (2 x 16 KB + 4 x 8 KB):
//                   what a pre-tiled global layout would allow,
// and a consumer thread hands the stage back as soon as it has landed.  Nothing reads the data.
// Prints GB/s per SM and chip-wide, for data that misses L2 (each CTA streams its own region) and
// for data that hits (all CTAs stream the same region).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/cuda/_bin/tma_rate_probe tests/cuda/tma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 65536;
constexpr long PITCH = 20096;          // bytes per plane row (157 stages of 128)
constexpr int N_STAGE_COLS = 157;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}" ::
            "r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                     "r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// mode 0: 2-D boxes (map128: box 128 B x 128 rows, map64: box 128 B x 64 rows); mode 1: 1-D bulk
__global__ void __launch_bounds__(64)
rate_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
            const uint8_t *flat, int mode, int shared_region, long rows_per_cta, int n_iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long cta = shared_region ? 0 : blockIdx.x;
    if (threadIdx.x == 0) {                       // producer
        for (int g = 0; g < n_iters; ++g) {
            const int s = g % STAGES;
            mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
            mbar_expect(&full[s], STAGE_BYTES);
            uint8_t *dst = smem + s * STAGE_BYTES;
            if (mode == 0) {
                const int col = (g % N_STAGE_COLS) * 128;
                const int row0 = (int)(cta * rows_per_cta) + ((g / N_STAGE_COLS) % 2) * 512;
                tma_2d(dst, &map128, col, row0, &full[s]);
                tma_2d(dst + 16384, &map128, col, row0 + 128, &full[s]);
                for (int a = 0; a < 4; ++a) tma_2d(dst + 32768 + a * 8192, &map64, col, row0 + 256 + a * 64, &full[s]);
            } else {
                const uint8_t *src = flat + (cta * rows_per_cta) * PITCH + (long)(g % (2 * N_STAGE_COLS)) * STAGE_BYTES;
                bulk_1d(dst, src, 16384, &full[s]);
                bulk_1d(dst + 16384, src + 16384, 16384, &full[s]);
                for (int a = 0; a < 4; ++a) bulk_1d(dst + 32768 + a * 8192, src + 32768 + a * 8192, 8192, &full[s]);
            }
        }
    } else if (threadIdx.x == 32) {               // consumer: hand the stage back at once
        for (int g = 0; g < n_iters; ++g) {
            const int s = g % STAGES;
            mbar_wait(&full[s], (g / STAGES) & 1);
            mbar_arrive(&empty[s]);
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long rows_per_cta = 1024;                       // 1024 rows x 20 096 B = 20.6 MB per CTA: 3 GB in all
    const long rows = rows_per_cta * sms;
    uint8_t *buf;
    CK(cudaMalloc(&buf, rows * PITCH));
    CK(cudaMemset(buf, 1, rows * PITCH));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    CUtensorMap map128, map64;
    cuuint64_t dims[2] = {(cuuint64_t)PITCH, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)PITCH};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t box128[2] = {128, 128}, box64[2] = {128, 64};
    if (enc(&map128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&map64, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box64, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed\n");
        return 2;
    }
    const int smem = STAGES * STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_iters = 2 * N_STAGE_COLS * 4;             // 1256 stages of 64 KB per CTA
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int shared_region = 0; shared_region < 2; ++shared_region)
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(a));
                rate_kernel<<<sms, 64, smem>>>(map128, map64, buf, mode, shared_region, rows_per_cta, n_iters);
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (rep && ms < best) best = ms;
            }
            const double bytes = (double)n_iters * STAGE_BYTES;
            printf("%-46s %-28s %6.3f ms  %6.1f GB/s per SM  %6.2f TB/s chip\n",
                   mode == 0 ? "2-D boxes, 128-byte rows at a 20 096-byte pitch" : "1-D bulk copies of contiguous 16 / 8 KB blocks",
                   shared_region ? "(all CTAs the same region)" : "(each CTA its own region)", best,
                   bytes / best / 1e6, bytes * sms / best / 1e9);
        }
    return 0;
}
