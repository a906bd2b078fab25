// Issue rate of tcgen05.mma kind::i8 by instruction shape (B200): how many SM cycles one MMA of
// K = 32 takes, for one CTA (cta_group::1, M = 128) and for a CTA pair (cta_group::2, M = 256),
// with N = 64, 128 and 256.  Operands are garbage in shared memory (canonical K-major layout, no
// swizzle; four different tiles in rotation), accumulators in tensor memory; every SM (pair of
// SMs) issues the same stream of MMAs and the elapsed time is divided by their number.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/_bin/umma_rate_probe tests/cuda/umma_rate_probe.cu
//
// Full int8 rate would be 8 192 MAC per clock and SM: 64 clocks for M = 128 (per SM), N = 128,
// K = 32.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    return d;
}

// PAIR = 0: one CTA, M = 128.  PAIR = 1: cluster of two, M = 256, each CTA holds its 128 rows of A
// and N/2 rows of B.
template <int N, int PAIR>
__global__ void __launch_bounds__(128) rate_kernel(int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;                       // 4 tiles of 128 rows x 32 B
    uint8_t *sb = smem + 4 * 4096;            // 4 tiles of up to 256 rows x 32 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 * 4096 + 4 * 8192);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < (4 * 4096 + 4 * 8192) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01020301u * (i & 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR)
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else
        __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const int brows = PAIR ? N / 2 : N;       // rows of B in THIS CTA's shared memory
    long long t0 = 0, t1 = 0;
    if (rank == 0 && tid == 0) {
        const uint32_t m = PAIR ? 256 : 128;
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                // one K step of 32 = two 16-byte chunks: LBO = chunk stride, SBO = 8-row group stride
                const uint64_t ad = make_desc(smem_u32(sa) + (q & 3) * 4096, 128 * 16, 128);
                const uint64_t bd = make_desc(smem_u32(sb) + (q & 3) * 8192, brows * 16, 128);
                const uint32_t d = tmem + (uint32_t)((q & 1) * N);       // two accumulators in rotation
                if (PAIR)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                                 "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                                 "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
            }
        }
        if (PAIR)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                             "r"(smem_u32(bar)),
                         "h"((uint16_t)3));
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::
                     "r"(smem_u32(bar)),
                 "r"(0));
    if (rank == 0 && tid == 0) {
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR)
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else
        __syncthreads();
    if (warp == 0) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

template <int N, int PAIR>
void run(int sms) {
    const int iters = 4000;
    const size_t smem = 4 * 4096 + 4 * 8192 + 64;
    long long *d_cycles;
    CK(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    CK(cudaMemset(d_cycles, 0, sizeof(long long) * sms));
    CK(cudaFuncSetAttribute(rate_kernel<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, rate_kernel<N, PAIR>, iters, d_cycles));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    long long h[256] = {0};
    CK(cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    const double n_mma = 8.0 * iters;
    const double mac_per_sm = 128.0 * N * 32;                  // per MMA and SM, either mode
    printf("%s M=%d N=%3d K=32: %8.1f clk per MMA (clock64 of CTA 0), %6.3f ms -> %6.0f MAC/clk/SM at 1.9 GHz, %5.2f int8 POP/s chip\n",
           PAIR ? "cta_group::2" : "cta_group::1", PAIR ? 256 : 128, N, (double)h[0] / n_mma, best,
           mac_per_sm * n_mma / (best * 1e-3) / 1.9e9, 2.0 * mac_per_sm * n_mma * sms / (best * 1e-3) / 1e15);
    CK(cudaFree(d_cycles));
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    run<64, 0>(sms);
    run<128, 0>(sms);
    run<256, 0>(sms);
    run<64, 1>(sms);
    run<128, 1>(sms);
    run<256, 1>(sms);
    return 0;
}
