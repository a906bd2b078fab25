// Hardware probe for the tcgen05 kind::i8 building blocks used by
// fastproject_b200/csrc/consistency_tc.cu.  Stand-alone (no torch):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_probe umma_i8_probe.cu -lcuda
//   ./umma_i8_probe
//
// Checks, against a CPU integer reference:
//   mode 0  A (s8) and B (u8) K-major, SWIZZLE_NONE canonical layout, filled by threads
//   mode 1  A in the SWIZZLE_128B K-major layout filled by threads (XOR pattern), B as mode 0
//   mode 2  A loaded by TMA (cuTensorMapEncodeTiled, 128B swizzle), B as mode 0
// and measures the issue rate of kind::i8 MMAs on all SMs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                       // descriptor version 1 (Blackwell)
    d |= (uint64_t)layout << 61;           // 0 none, 2 = 128B swizzle
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int M = 128, N = 128, KB = 128;   // one stage: K = 128 bytes = 4 MMA k-steps

struct __align__(1024) Smem {
    uint8_t a[M * KB];       // 16 KB
    uint8_t b[N * KB];       // 16 KB
    uint64_t bar_tma, bar_mma;
    uint32_t tmem_base;
};

// A: [M][KB] int8 row-major in global, B: [N][KB] uint8 row-major, D: [M][N] int32
__global__ void __launch_bounds__(128) probe_kernel(const int8_t *A, const uint8_t *B, int32_t *D, int mode,
                                                    const __grid_constant__ CUtensorMap tmap, int a_signed,
                                                    int b_signed) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sm.bar_tma, 1);
        mbar_init(&sm.bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // B: no swizzle, [chunk c][row][16 B]
    for (int idx = tid; idx < N * (KB / 16); idx += 128) {
        const int row = idx % N, c = idx / N;
        const uint4 v = *reinterpret_cast<const uint4 *>(B + row * KB + c * 16);
        *reinterpret_cast<uint4 *>(sm.b + (c * N + row) * 16) = v;
    }
    if (mode == 0) {
        for (int idx = tid; idx < M * (KB / 16); idx += 128) {
            const int row = idx % M, c = idx / M;
            const uint4 v = *reinterpret_cast<const uint4 *>(A + row * KB + c * 16);
            *reinterpret_cast<uint4 *>(sm.a + (c * M + row) * 16) = v;
        }
    } else if (mode == 1) {
        // 128B swizzle: row r at r*128, 16-byte chunk c stored at chunk (c ^ (r & 7))
        for (int idx = tid; idx < M * (KB / 16); idx += 128) {
            const int row = idx / 8, c = idx % 8;
            const uint4 v = *reinterpret_cast<const uint4 *>(A + row * KB + c * 16);
            *reinterpret_cast<uint4 *>(sm.a + row * 128 + ((c ^ (row & 7)) * 16)) = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_base;
    if (mode == 2 && tid == 0) {
        mbar_expect_tx(&sm.bar_tma, M * KB);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                smem_u32(sm.a)),
            "l"(&tmap), "r"(0), "r"(0), "r"(smem_u32(&sm.bar_tma))
            : "memory");
    }
    if (tid == 0) {
        if (mode == 2) mbar_wait(&sm.bar_tma, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc_i8(M, N, a_signed, b_signed);
        for (int ks = 0; ks < KB / 32; ++ks) {
            uint64_t ad, bd;
            if (mode == 0)
                ad = make_desc(smem_u32(sm.a) + ks * 2 * M * 16, M * 16, 128, 0);
            else
                ad = make_desc(smem_u32(sm.a) + ks * 32, 16, 1024, 2);
            bd = make_desc(smem_u32(sm.b) + ks * 2 * N * 16, N * 16, 128, 0);
            umma_i8(tmem, ad, bd, idesc, ks > 0);
        }
        umma_commit(&sm.bar_mma);
    }
    mbar_wait(&sm.bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[(warp * 32 + (tid & 31)) * N + c0 + i] = (int32_t)v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// issue-rate kernel: every CTA issues iters x 8 MMAs (M=128, N=NN, K=32) on garbage operands
template <int NN>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int32_t *sink) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(raw + 96 * 1024);
    uint32_t *tbase = reinterpret_cast<uint32_t *>(raw + 96 * 1024 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(raw)[i] = 0x01010101u * (i & 3);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tbase;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(128, NN, 1, 0);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint64_t ad = make_desc(smem_u32(raw) + (q & 3) * 4096, 2048, 128, 0);
                const uint64_t bd = make_desc(smem_u32(raw) + 32768 + (q & 3) * 8192, NN * 16, 128, 0);
                umma_i8(tmem + (q & 1) * NN, ad, bd, idesc, 1);
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    if (v[0] == 0x7fffffff) sink[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<int8_t> hA(M * KB);
    std::vector<uint8_t> hB(N * KB);
    srand(7);
    for (auto &v : hA) v = (int8_t)(rand() % 256 - 128);
    for (auto &v : hB) v = (uint8_t)(rand() % 256);
    std::vector<int32_t> ref(M * N), got(M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            int32_t s = 0;
            for (int k = 0; k < KB; ++k) s += (int32_t)hA[m * KB + k] * (int32_t)hB[n * KB + k];
            ref[m * N + n] = s;
        }
    int8_t *dA;
    uint8_t *dB;
    int32_t *dD;
    CK(cudaMalloc(&dA, hA.size()));
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dD, ref.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        cuuint64_t dims[2] = {(cuuint64_t)KB, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)KB};
        cuuint32_t box[2] = {128, (cuuint32_t)M};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
    }
    const size_t smem = sizeof(Smem) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int fails = 0;
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(dD, 0xff, ref.size() * 4));
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, tmap, 1, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e));
            return 3;
        }
        CK(cudaMemcpy(got.data(), dD, ref.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (size_t i = 0; i < ref.size(); ++i) bad += got[i] != ref[i];
        printf("mode %d: %d / %d mismatches (D[0]=%d ref %d, D[last]=%d ref %d)\n", mode, bad, (int)ref.size(),
               got[0], ref[0], got[M * N - 1], ref[M * N - 1]);
        fails += bad != 0;
    }
    // exactness at the int32 limit: all A = -128, all B = 255 -> -32640 per term
    {
        std::vector<int8_t> a2(M * KB, (int8_t)-128);
        std::vector<uint8_t> b2(N * KB, (uint8_t)255);
        CK(cudaMemcpy(dA, a2.data(), a2.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, b2.data(), b2.size(), cudaMemcpyHostToDevice));
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 0, tmap, 1, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), dD, ref.size() * 4, cudaMemcpyDeviceToHost));
        printf("extreme digits: D[0]=%d expected %d\n", got[0], -128 * 255 * KB);
        fails += got[0] != -128 * 255 * KB;
    }
    // issue rate
    int32_t *sink;
    CK(cudaMalloc(&sink, 4));
    int dev = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem2 = 97 * 1024;
    CK(cudaFuncSetAttribute(rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CK(cudaFuncSetAttribute(rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int nn = 128; nn <= 256; nn += 128) {
        for (int rep = 0; rep < 2; ++rep) {
            const int iters = 20000;
            CK(cudaEventRecord(e0));
            if (nn == 128)
                rate_kernel<128><<<sms, 128, smem2>>>(iters, sink);
            else
                rate_kernel<256><<<sms, 128, smem2>>>(iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double macs = (double)sms * iters * 8.0 * 128.0 * nn * 32.0;
            printf("rate N=%d: %.3f ms, %.1f int8 TOP/s (2*MAC), %.1f MMA-cycles@1.9GHz per instr\n", nn, ms,
                   2.0 * macs / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 / (iters * 8.0));
        }
    }
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails ? 1 : 0;
}
