// Signature scoring: K signatures x N cells in one launch.
//
// Replaces the bodies of naive/weighted/imputed/nonzero_eval_signature
// (/root/reference/FastProject/SigScoreMethods.py:18-162) and the Python loop
// of calculate_sig_scores (Signatures.py:669-679).
//
// Bit-exactness contract: the reference gathers the signature's rows in
// ascending gene order and reduces with ndarray.sum(axis=0), which NumPy
// evaluates row after row (sequential in g).  Every product below is a
// separately rounded __dmul_rn and every accumulation a __dadd_rn in ascending
// gene order, so no FMA contraction or reassociation can change a bit.
#include <stdlib.h>
#include <string.h>

#include "row_tree.cuh"

namespace {

// Mapping: one warp = one signature x 32 consecutive cells; lanes read 256
// contiguous bytes of a gene row (two full 128-byte lines) per operand.
// Launch order: blockIdx.x runs over the signature tiles of ONE slab of 32
// cells, blockIdx.y over the slabs, so the CTAs in flight all gather from the
// same few G x 32-cell slabs of X and W (7.7 MB each at G = 15 000), which are
// served from L2 to the ~15 signatures that touch each gene row.
// (Round-1 profile of the transposed order: 73.9 GB of DRAM reads for 4.8 GB of
// input, L2 hit rate 0.7 %.)
//
// Memory-level parallelism: a thread first loads the indices of kDepth genes,
// then all their operands, and only then accumulates -- in ascending gene
// order, every operation separately rounded, exactly as a one-gene-at-a-time
// loop would (bit-identical results).  The first version issued the loads of
// one gene, waited, accumulated, and only then addressed the next gene
// (volatile loads, 32 registers per thread): ~1 300 clk per gene and warp,
// latency bound at 7 TB/s of L2 gather; eight genes in flight at 64 registers
// and 1 024 threads per SM reach 13 TB/s, the L2 -> SM limit of the part
// (8.5 -> 4.7 ms at 3 000 signatures x 20 000 cells, tests/cuda/score_profile.py).
constexpr int kCellsPerBlock = 32;    // blockDim.x: one warp wide -> slab of G x 32 cells
constexpr int kSigsPerBlock = 16;     // blockDim.y
constexpr int kDepth = 8;             // genes in flight per thread
constexpr int kMinBlocks = 2;         // CTAs per SM the register budget is set for (64 registers)

__device__ __forceinline__ double ldg_stream(const double *p, uint64_t policy) {
    // read-only, keep out of L1 (each value is used once by this thread) but LAST to leave L2:
    // the same gene row is gathered by ~nnz/G other signatures of this 32-cell slab within
    // microseconds (ncu, round 2: without the hint the slab was re-fetched from HBM 3.7 times,
    // 17.7 GB of DRAM reads for 4.8 GB of input).  Not volatile, so the compiler may hoist the
    // loads of later genes above the arithmetic of earlier ones.
    double v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

// one gene's contribution, the reference's operation order
template <int METHOD>
__device__ __forceinline__ void accumulate(double x, double w, double z, double mu, double s, double &num,
                                           double &den, double &aux) {
    if (METHOD == FP_METHOD_NAIVE) {
        // pdata = data * sig * ones ; norm = |sig| * ones   (:38-41)
        num = __dadd_rn(num, __dmul_rn(__dmul_rn(x, s), 1.0));
        den = __dadd_rn(den, 1.0);
    } else if (METHOD == FP_METHOD_WEIGHTED) {
        // pdata = data * sig * weights ; norm = |sig| * weights   (:70-73)
        num = __dadd_rn(num, __dmul_rn(__dmul_rn(x, s), w));
        den = __dadd_rn(den, w);
    } else if (METHOD == FP_METHOD_NONZERO) {
        // (data * not_zeros) * sig ; norm = sum(not_zeros)   (:152-153)
        const double nz = __dsub_rn(1.0, z);
        num = __dadd_rn(num, __dmul_rn(__dmul_rn(x, nz), s));
        den = __dadd_rn(den, nz);
    } else {  // FP_METHOD_IMPUTED (:105-121)
        const double nz = __dsub_rn(1.0, z);
        num = __dadd_rn(num, __dmul_rn(__dmul_rn(x, nz), s));
        const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, w), mu), s), z);
        const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(w, x), s), z);
        aux = __dadd_rn(aux, __dadd_rn(t1, t2));
    }
}

template <int METHOD>
__global__ void __launch_bounds__(kCellsPerBlock *kSigsPerBlock, kMinBlocks)
score_gather_kernel(const double *__restrict__ X, const double *__restrict__ W,
                    const double *__restrict__ Z, long N, long ld,
                    const int32_t *__restrict__ sig_ptr, const int32_t *__restrict__ sig_gene,
                    const int8_t *__restrict__ sig_sign, long K,
                    const double *__restrict__ gene_mu, double *__restrict__ out, long ld_out) {
    constexpr bool kUseW = METHOD == FP_METHOD_WEIGHTED || METHOD == FP_METHOD_IMPUTED;
    constexpr bool kUseZ = METHOD == FP_METHOD_NONZERO || METHOD == FP_METHOD_IMPUTED;
    constexpr bool kUseMu = METHOD == FP_METHOD_IMPUTED;
    // the imputed method carries three operands + mu per gene: half the depth keeps it in registers
    constexpr int D = METHOD == FP_METHOD_IMPUTED ? kDepth / 2 : kDepth;
    const uint64_t keep = l2_evict_last_policy();
    for (long slab = blockIdx.y; slab * kCellsPerBlock < N; slab += gridDim.y) {
    const long cell = slab * kCellsPerBlock + threadIdx.x;
    const bool live = cell < N;
    const long c = live ? cell : (N - 1);
    for (long k = (long)blockIdx.x * kSigsPerBlock + threadIdx.y; k < K;
         k += (long)gridDim.x * kSigsPerBlock) {
        const int beg = sig_ptr[k], end = sig_ptr[k + 1];
        double num = 0.0, den = 0.0, aux = 0.0;
        int t = beg;
        for (; t + D <= end; t += D) {
            int g[D];
            double x[D], w[D], z[D], mu[D], s[D];
#pragma unroll
            for (int e = 0; e < D; ++e) g[e] = sig_gene[t + e];
#pragma unroll
            for (int e = 0; e < D; ++e) {
                const long o = (long)g[e] * ld + c;
                x[e] = ldg_stream(X + o, keep);
                w[e] = kUseW ? ldg_stream(W + o, keep) : 0.0;
                z[e] = kUseZ ? ldg_stream(Z + o, keep) : 0.0;
            }
#pragma unroll
            for (int e = 0; e < D; ++e) {
                s[e] = (double)sig_sign[t + e];
                mu[e] = kUseMu ? gene_mu[g[e]] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < D; ++e) accumulate<METHOD>(x[e], w[e], z[e], mu[e], s[e], num, den, aux);
        }
        for (; t < end; ++t) {
            const int g = sig_gene[t];
            const long o = (long)g * ld + c;
            const double x = ldg_stream(X + o, keep);
            const double w = kUseW ? ldg_stream(W + o, keep) : 0.0;
            const double z = kUseZ ? ldg_stream(Z + o, keep) : 0.0;
            accumulate<METHOD>(x, w, z, kUseMu ? gene_mu[g] : 0.0, (double)sig_sign[t], num, den, aux);
        }
        double score;
        if (METHOD == FP_METHOD_NAIVE) {
            score = __ddiv_rn(num, den);
        } else if (METHOD == FP_METHOD_IMPUTED) {
            score = __ddiv_rn(__dadd_rn(num, aux), (double)(end - beg));
        } else {
            if (den == 0.0) den = 1.0;           // norm_factor[norm_factor == 0] = 1.0
            score = __ddiv_rn(num, den);
        }
        if (live) out[k * ld_out + cell] = score;
    }
    }
}

// mu_g over all cells, NumPy pairwise order (SigScoreMethods.py:109-111).
struct DetectedProduct {
    const double *x, *z;
    __device__ double operator()(long i) const { return __dmul_rn(x[i], __dsub_rn(1.0, z[i])); }
};
struct DetectedCount {
    const double *z;
    __device__ double operator()(long i) const { return __dsub_rn(1.0, z[i]); }
};

// One CTA per gene (persistent over the genes): the CTA enumerates the pairwise tree of a row of
// N cells once into shared memory, then, per gene, eight lanes sum each <= 128-element leaf
// (lane k owns NumPy's k-th interleaved accumulator: coalesced 64-byte loads, rowtree::leaf_sums)
// and one thread folds the leaf sums with the tree's post-order program -- NumPy's
// add.reduce(axis=1) bit for bit, as normalize.cu and weights.cu do it.
constexpr int MU_THREADS = 256;

__global__ void __launch_bounds__(MU_THREADS)
gene_detected_mean_kernel(const double *__restrict__ X, const double *__restrict__ Z, long G, long N, long ld,
                          int max_leaves, double *__restrict__ out_mu) {
    extern __shared__ __align__(16) unsigned char mu_smem[];
    double *leaf_sum = reinterpret_cast<double *>(mu_smem);                               // max_leaves
    fp::rowtree::Leaf *leaves = reinterpret_cast<fp::rowtree::Leaf *>(leaf_sum + max_leaves);
    int *prog = reinterpret_cast<int *>(leaves + max_leaves);                              // 2 * max_leaves
    __shared__ double stack[fp::rowtree::MAX_DEPTH];
    __shared__ int counts[2];
    __shared__ double num_s;
    if (threadIdx.x == 0) fp::rowtree::build_tree_device(0, N, leaves, &counts[0], prog, &counts[1]);
    __syncthreads();
    const int n_leaves = counts[0], n_prog = counts[1];
    for (long g = blockIdx.x; g < G; g += gridDim.x) {
        const double *x = X + g * ld, *z = Z + g * ld;
        fp::rowtree::leaf_sums(DetectedProduct{x, z}, leaves, n_leaves, leaf_sum);
        __syncthreads();
        if (threadIdx.x == 0) num_s = fp::rowtree::fold_tree(leaf_sum, prog, n_prog, stack);
        __syncthreads();
        fp::rowtree::leaf_sums(DetectedCount{z}, leaves, n_leaves, leaf_sum);
        __syncthreads();
        if (threadIdx.x == 0) {
            const double den = fp::rowtree::fold_tree(leaf_sum, prog, n_prog, stack);
            double mu = __ddiv_rn(num_s, den);
            if (mu != mu) mu = 0.0;                  // mu[np.isnan(mu)] = 0.0
            out_mu[g] = mu;
        }
        __syncthreads();
    }
}

// rows too long for the tree to fit in shared memory (> ~1.2 M cells): one thread per gene
__global__ void gene_detected_mean_long_kernel(const double *__restrict__ X, const double *__restrict__ Z,
                                               long G, long N, long ld, double *__restrict__ out_mu) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    DetectedProduct prod{X + g * ld, Z + g * ld};
    DetectedCount cnt{Z + g * ld};
    const double num = fp::np_pairwise_sum(prod, 0, N);
    const double den = fp::np_pairwise_sum(cnt, 0, N);
    double mu = __ddiv_rn(num, den);
    if (mu != mu) mu = 0.0;
    out_mu[g] = mu;
}

}  // namespace

extern "C" int fp_gene_detected_mean(const double *X, const double *Z, int64_t G, int64_t N,
                                     int64_t ld, double *out_mu, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(X && Z && out_mu, "fp_gene_detected_mean: null pointer");
    FP_CHECK_ARG(G > 0 && N > 0 && ld >= N, "fp_gene_detected_mean: bad shape G=%ld N=%ld ld=%ld",
                 (long)G, (long)N, (long)ld);
    // leaves of the pairwise tree of a row of N cells (the host enumeration only sizes the
    // shared-memory tables; every CTA builds its own copy on the device)
    long max_leaves = 1;
    if (N < (1L << 31)) {
        std::vector<fp::rowtree::Leaf> leaves;
        std::vector<int> prog;
        int depth = 0;
        fp::rowtree::build_tree(0, N, leaves, prog, 1, depth);
        max_leaves = (long)leaves.size();
    }
    const size_t smem = (size_t)max_leaves * (sizeof(double) + sizeof(fp::rowtree::Leaf) + 2 * sizeof(int));
    if (N < (1L << 31) && smem <= 200 * 1024) {
        static size_t smem_set = 0;
        if (smem > 40 * 1024 && smem > smem_set) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(gene_detected_mean_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
        }
        long blocks = G;
        const long cap = (long)fp::sm_count() * 8;
        if (blocks > cap) blocks = cap;
        gene_detected_mean_kernel<<<(unsigned)blocks, MU_THREADS, smem, fp::as_stream(stream)>>>(
            X, Z, G, N, ld, (int)max_leaves, out_mu);
        FP_CHECK_LAUNCH("gene_detected_mean_kernel");
        return FP_OK;
    }
    const int threads = 32;
    gene_detected_mean_long_kernel<<<(unsigned)((G + threads - 1) / threads), threads, 0,
                                     fp::as_stream(stream)>>>(X, Z, G, N, ld, out_mu);
    FP_CHECK_LAUNCH("gene_detected_mean_long_kernel");
    return FP_OK;
}

// score_staged.cu
size_t fp_score_staged_workspace_bytes(long G, long K, long nnz);
bool fp_score_staged_applies(const double *X, const double *W, const double *Z, long G, long N, long ld, long K,
                             long nnz, bool forced);
int fp_score_staged(int method, const double *X, const double *W, const double *Z, long G, long N, long ld,
                    const int32_t *sig_ptr, const int32_t *sig_gene, const int8_t *sig_sign, long K, long nnz,
                    const double *gene_mu, double *out, long ld_out, void *workspace, cudaStream_t st);

extern "C" size_t fp_score_workspace_bytes(int64_t G, int64_t K, int64_t nnz) {
    return fp_score_staged_workspace_bytes((long)G, (long)K, (long)nnz);
}

extern "C" int fp_score_signatures(int method, const double *X, const double *W, const double *Z,
                                   int64_t G, int64_t N, int64_t ld, const int32_t *sig_ptr,
                                   const int32_t *sig_gene, const int8_t *sig_sign, int64_t K, int64_t nnz,
                                   const double *gene_mu, double *out_scores, int64_t ld_out,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(method >= FP_METHOD_NAIVE && method <= FP_METHOD_NONZERO,
                 "fp_score_signatures: unknown method %d", method);
    FP_CHECK_ARG(X && sig_ptr && sig_gene && sig_sign && out_scores,
                 "fp_score_signatures: null pointer");
    FP_CHECK_ARG(G > 0 && N > 0 && ld >= N && ld_out >= N && K >= 0 && nnz >= 0,
                 "fp_score_signatures: bad shape G=%ld N=%ld ld=%ld K=%ld ld_out=%ld", (long)G,
                 (long)N, (long)ld, (long)K, (long)ld_out);
    FP_CHECK_ARG(method == FP_METHOD_NAIVE || method == FP_METHOD_NONZERO || W,
                 "fp_score_signatures: method %d needs W", method);
    FP_CHECK_ARG(method == FP_METHOD_NAIVE || method == FP_METHOD_WEIGHTED || Z,
                 "fp_score_signatures: method %d needs Z", method);
    FP_CHECK_ARG(method != FP_METHOD_IMPUTED || gene_mu, "fp_score_signatures: imputed needs gene_mu");
    if (K == 0) return FP_OK;
    // With the membership count and a workspace the TMA-staged kernel is used when it moves fewer
    // bytes through L2 than the gathers would (score_staged.cu); FASTPROJECT_B200_SCORE = gather |
    // staged forces one of them (tests compare the two bit for bit)
    if (nnz > 0 && workspace && workspace_bytes >= fp_score_staged_workspace_bytes((long)G, (long)K, (long)nnz)) {
        const char *force = getenv("FASTPROJECT_B200_SCORE");
        const bool want_gather = force && !strcmp(force, "gather");
        const bool want_staged = force && !strcmp(force, "staged");
        const bool use_w = method == FP_METHOD_WEIGHTED || method == FP_METHOD_IMPUTED;
        const bool use_z = method == FP_METHOD_NONZERO || method == FP_METHOD_IMPUTED;
        if (!want_gather && fp_score_staged_applies(X, use_w ? W : nullptr, use_z ? Z : nullptr, (long)G, (long)N,
                                                    (long)ld, (long)K, (long)nnz, want_staged))
            return fp_score_staged(method, X, W, Z, (long)G, (long)N, (long)ld, sig_ptr, sig_gene, sig_sign, (long)K,
                                   (long)nnz, gene_mu, out_scores, (long)ld_out, workspace, fp::as_stream(stream));
    }
    dim3 block(kCellsPerBlock, kSigsPerBlock);
    long cell_tiles = (N + kCellsPerBlock - 1) / kCellsPerBlock;
    if (cell_tiles > 65535) cell_tiles = 65535;
    const long sig_tiles = (K + kSigsPerBlock - 1) / kSigsPerBlock;
    dim3 grid((unsigned)sig_tiles, (unsigned)cell_tiles);
    cudaStream_t st = fp::as_stream(stream);
#define FP_LAUNCH_SCORE(M)                                                                    \
    score_gather_kernel<M><<<grid, block, 0, st>>>(X, W, Z, N, ld, sig_ptr, sig_gene, sig_sign, \
                                                   K, gene_mu, out_scores, ld_out)
    switch (method) {
        case FP_METHOD_NAIVE: FP_LAUNCH_SCORE(FP_METHOD_NAIVE); break;
        case FP_METHOD_WEIGHTED: FP_LAUNCH_SCORE(FP_METHOD_WEIGHTED); break;
        case FP_METHOD_IMPUTED: FP_LAUNCH_SCORE(FP_METHOD_IMPUTED); break;
        default: FP_LAUNCH_SCORE(FP_METHOD_NONZERO); break;
    }
#undef FP_LAUNCH_SCORE
    FP_CHECK_LAUNCH("score_gather_kernel");
    return FP_OK;
}
