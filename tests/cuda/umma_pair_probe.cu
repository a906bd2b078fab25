// Probe of tcgen05.mma cta_group::2 kind::i8 operand placement (B200).
// Cluster of 2 CTAs; C[256 x N] = A[256 x K] * B[N x K]^T, K = 64.
// Each CTA puts 128 rows of A (SWIZZLE_NONE K-major) in its smem; B placement is varied:
//   variant 0: CTA r holds B rows [r*N/2, (r+1)*N/2) at the descriptor address
//   variant 1: both CTAs hold ALL N rows of B at the descriptor address
// Output: per variant, mismatch count of each CTA's TMEM against the reference, and, to
// help decode other conventions, which B row each output column actually matches.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

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
constexpr int KB = 64;

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_kernel(const int8_t *A, const uint8_t *B, int32_t *D, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;                 // 128 rows x 64 B, [chunk][row][16]
    uint8_t *sb = smem + 128 * KB;      // up to 256 rows x 64 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 128 * KB + 256 * KB);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    for (int idx = tid; idx < 128 * (KB / 16); idx += 128) {
        const int row = idx % 128, c = idx / 128;
        *reinterpret_cast<uint4 *>(sa + (c * 128 + row) * 16) =
            *reinterpret_cast<const uint4 *>(A + (rank * 128 + row) * KB + c * 16);
    }
    const int brows = variant == 0 ? N / 2 : N;
    const int boff = variant == 0 ? rank * (N / 2) : 0;
    for (int idx = tid; idx < brows * (KB / 16); idx += 128) {
        const int row = idx % brows, c = idx / brows;
        *reinterpret_cast<uint4 *>(sb + (c * brows + row) * 16) =
            *reinterpret_cast<const uint4 *>(B + (boff + row) * KB + c * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int ks = 0; ks < KB / 32; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * brows * 16, brows * 16, 128);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u));
        }
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                smem_u32(bar)),
            "h"((uint16_t)3));
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(0));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[(rank * 128 + warp * 32 + (tid & 31)) * N + c0 + i] = (int32_t)v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int N>
int run() {
    std::vector<int8_t> hA(256 * KB);
    std::vector<uint8_t> hB(256 * KB);
    srand(11);
    for (auto &v : hA) v = (int8_t)(rand() % 256 - 128);
    for (auto &v : hB) v = (uint8_t)(rand() % 256);
    std::vector<int32_t> ref(256 * 256), got(256 * N);
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < 256; ++n) {
            int32_t s = 0;
            for (int k = 0; k < KB; ++k) s += (int32_t)hA[m * KB + k] * (int32_t)hB[n * KB + k];
            ref[m * 256 + n] = s;
        }
    int8_t *dA; uint8_t *dB; int32_t *dD;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, got.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    const size_t smem = 128 * KB + 256 * KB + 64;
    CK(cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int fails = 0;
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(dD, 0xff, got.size() * 4));
        pair_kernel<N><<<2, 128, smem>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d variant %d: CUDA error %s\n", N, variant, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < N; ++n) bad += got[m * N + n] != ref[m * 256 + n];
        printf("N=%d variant %d: %d / %d mismatches.", N, variant, bad, 256 * N);
        // decode: for output column n (row 0 and row 128), which B row does it equal?
        for (int m : {0, 128}) {
            printf("  [row %d: cols", m);
            for (int n : {0, 1, N / 2 - 1, N / 2, N - 1}) {
                int match = -1;
                for (int q = 0; q < 256; ++q) if (got[m * N + n] == ref[m * 256 + q]) { match = q; break; }
                printf(" %d->B%d", n, match);
            }
            printf("]");
        }
        printf("\n");
        fails += bad != 0;
    }
    return fails;
}

int main() {
    int f = run<128>();
    f += run<256>();
    f += run<64>();
    printf("done (%d variant(s) mismatched)\n", f);
    return 0;
}
