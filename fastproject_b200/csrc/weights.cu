// False-negative weights of the expression matrix -- the producer of the W matrix that the
// scoring kernels read.  Replaces
// /root/reference/FastProject/Transforms.py:266-316 (compute_weights) with the logistic
// false-negative curve of create_false_neg_map (:87-88) as fit function, and :341-381
// (estimate_non_detection with DO_GMM = False: p_nd = [x == 0]) when no p_nd is passed.
//
//   p_d      = 1 - p_nd
//   mu_g     = sum_n(x * p_d) / sum_n(p_d)                    (mean of the detected values)
//   fn_gn    = L_n + S_n / (1 + exp((mu_g - x0_n) * a_n))     (false-negative rate)
//   pd_e     = 1 - fn
//   pnd_g    = mean_n(p_nd);  pe_g = (1 - pnd_g) / mean_n(pd_e);  pnd_g == 0 -> 1 / N
//   pne_nd   = clip(1 - (1 - pd_e) * pe / pnd, 0, 1)
//   W        = pne_nd * p_nd + p_d
//
// One CTA per gene, three sweeps over the row (the second and third hit L2; the second parks
// 1 - fn in the output row so that exp and the division run once per element): the row sums
// follow NumPy's pairwise order (row_tree.cuh) and every operation is rounded separately in
// the reference's order, so the only difference from the reference is the last-bit freedom of
// exp() (CUDA's and NumPy's are both within 1 ulp but not identical): weights agree to
// ~1e-15, tested at 1e-12.  NaN propagates as in the reference (a gene that is never
// detected has mu = 0/0; the reference's line 305 compares instead of assigning).
#include "row_tree.cuh"

namespace {

using fp::rowtree::Leaf;
using fp::rowtree::Node;
using fp::rowtree::fold_levels;
using fp::rowtree::build_tree;
using fp::rowtree::MAX_DEPTH;
using fp::rowtree::align_up;

constexpr int W_THREADS = 256;

__device__ __forceinline__ double p_nd_at(const double *x, const double *pnd_row, long i) {
    return pnd_row ? pnd_row[i] : (x[i] == 0.0 ? 1.0 : 0.0);
}

struct DetectedValue {       // x * p_d
    const double *x, *pnd;
    __device__ double operator()(long i) const { return __dmul_rn(x[i], __dsub_rn(1.0, p_nd_at(x, pnd, i))); }
};
struct Detected {            // p_d
    const double *x, *pnd;
    __device__ double operator()(long i) const { return __dsub_rn(1.0, p_nd_at(x, pnd, i)); }
};
struct NotDetected {         // p_nd
    const double *x, *pnd;
    __device__ double operator()(long i) const { return p_nd_at(x, pnd, i); }
};

__device__ __forceinline__ double fn_rate(double mu, const double *__restrict__ params, long ldp, long i) {
    const double x0 = params[i], a = params[ldp + i], L = params[2 * ldp + i], S = params[3 * ldp + i];
    const double e = exp(__dmul_rn(__dsub_rn(mu, x0), a));
    return __dadd_rn(L, __ddiv_rn(S, __dadd_rn(1.0, e)));
}
struct ExpressedDetect {     // pd_e = 1 - fn; also parked in the output row for the last sweep
    double mu;
    const double *params;
    long ldp;
    double *park;
    __device__ double operator()(long i) const {
        const double v = __dsub_rn(1.0, fn_rate(mu, params, ldp, i));
        park[i] = v;         // (the n % 8 tail of a leaf is evaluated by all eight lanes: same value)
        return v;
    }
};

__global__ void __launch_bounds__(W_THREADS)
compute_weights_kernel(const double *__restrict__ X, const double *__restrict__ P_nd, const double *__restrict__ params,
                       long ldp, long G, long N, long ld, long ld_pnd, double *__restrict__ out, long ld_out,
                       const Leaf *__restrict__ leaves, int n_leaves, const Node *__restrict__ nodes, int height) {
    extern __shared__ double val[];                      // 2 * n_leaves: one tree at a time
    for (long g = blockIdx.x; g < G; g += gridDim.x) {
        const double *row = X + g * ld;
        const double *prow = P_nd ? P_nd + g * ld_pnd : nullptr;
        fp::rowtree::leaf_sums(DetectedValue{row, prow}, leaves, n_leaves, val);
        __syncthreads();
        const double sxp = fold_levels(val, nodes, n_leaves, height);
        fp::rowtree::leaf_sums(Detected{row, prow}, leaves, n_leaves, val);
        __syncthreads();
        const double sp = fold_levels(val, nodes, n_leaves, height);
        fp::rowtree::leaf_sums(NotDetected{row, prow}, leaves, n_leaves, val);
        __syncthreads();
        const double snd = fold_levels(val, nodes, n_leaves, height);
        const double mu = __ddiv_rn(sxp, sp);
        double pnd = __ddiv_rn(snd, (double)N);
        double *orow = out + g * ld_out;
        // the exp and the division of the false-negative curve are evaluated once per element:
        // this sweep parks 1 - fn in the output row, the last sweep turns it into the weight
        fp::rowtree::leaf_sums(ExpressedDetect{mu, params, ldp, orow}, leaves, n_leaves, val);
        __syncthreads();
        const double mean_pde = __ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N);
        const double pe = __ddiv_rn(__dsub_rn(1.0, pnd), mean_pde);
        if (pnd == 0.0) pnd = __ddiv_rn(1.0, (double)N);           // "for stability" (:307)
        for (long c = threadIdx.x; c < N; c += blockDim.x) {
            const double p_nd = p_nd_at(row, prow, c);
            const double p_d = __dsub_rn(1.0, p_nd);
            const double pd_e = orow[c];
            // 1 - (1 - pd_e) * pe / pnd, evaluated left to right
            double v = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(__dsub_rn(1.0, pd_e), pe), pnd));
            if (v < 0.0) v = 0.0;
            if (v > 1.0) v = 1.0;
            orow[c] = __dadd_rn(__dmul_rn(v, p_nd), p_d);
        }
        __syncthreads();
    }
}

// The pipeline's case (p_nd is the zero mask of the data: estimate_non_detection with DO_GMM =
// False) for rows that fit in shared memory.  The row is staged once (row_tree.cuh); with a 0/1
// mask x * p_d is x itself and sum(p_d) / sum(p_nd) are counts -- integers, exact in any order, so
// a ballot count replaces two of the three tree sums; the sweep that evaluates the curve writes
// 1 - fn over the staged row (the zero mask moves to a bitmap first), and the last sweep reads
// it back from shared memory: DRAM sees one read and one write of the matrix.
constexpr int STAGED_THREADS = 1024;

struct StagedValue {
    const double *row;
    __device__ double operator()(long i) const { return row[i]; }
};
struct StagedExpressedDetect {   // pd_e = 1 - fn, left in place of x
    double mu;
    const double *params;
    long ldp;
    double *row;
    __device__ double operator()(long i) const {
        const double v = __dsub_rn(1.0, fn_rate(mu, params, ldp, i));
        row[i] = v;
        return v;
    }
};

__global__ void __launch_bounds__(STAGED_THREADS)
compute_weights_staged_kernel(const double *__restrict__ X, const double *__restrict__ params, long ldp, long G, long N,
                              long ld, double *__restrict__ out, long ld_out, const Leaf *__restrict__ leaves,
                              int n_leaves, const Node *__restrict__ nodes, int height) {
    extern __shared__ __align__(16) double sm[];         // row | 2 * n_leaves | zero bitmap
    __shared__ uint64_t bar;
    __shared__ unsigned n_zero;
    const long n_even = (N + 1) & ~1L;
    double *row = sm, *val = sm + n_even;
    uint32_t *zero_bits = reinterpret_cast<uint32_t *>(val + 2 * n_leaves);
    const long n_bulk = N & ~1L;
    if (threadIdx.x == 0) {
        fp::rowtree::stage_init(&bar);
        if (n_bulk && blockIdx.x < G) fp::rowtree::stage_row_async(row, X + blockIdx.x * ld, (uint32_t)(n_bulk * 8), &bar);
    }
    __syncthreads();
    uint32_t parity = 0;
    const long n_words = (N + 31) >> 5;
    for (long g = blockIdx.x; g < G; g += gridDim.x) {
        const double *src = X + g * ld;
        const long next = g + gridDim.x;
        if (threadIdx.x == 32 && next < G && n_bulk) fp::rowtree::prefetch_row_l2(X + next * ld, (uint32_t)(n_bulk * 8));
        if (threadIdx.x == 0) n_zero = 0;
        if (n_bulk) {
            fp::rowtree::stage_wait(&bar, parity);
            parity ^= 1;
        }
        if ((N & 1) && threadIdx.x == 0) row[N - 1] = src[N - 1];
        __syncthreads();
        // zero mask as a bitmap + its population (warp w owns words w, w + 32, ...)
        unsigned zeros = 0;
        for (long w = threadIdx.x >> 5; w < n_words; w += STAGED_THREADS / 32) {
            const long c = (w << 5) + (threadIdx.x & 31);
            const unsigned bits = __ballot_sync(0xffffffffu, c < N && row[c] == 0.0);
            if ((threadIdx.x & 31) == 0) {
                zero_bits[w] = bits;
                zeros += __popc(bits);
            }
        }
        if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(&n_zero, zeros);
        fp::rowtree::leaf_sums(StagedValue{row}, leaves, n_leaves, val);
        __syncthreads();
        const double sxp = fold_levels(val, nodes, n_leaves, height);
        const double snd = (double)n_zero, sp = (double)(N - (long)n_zero);
        const double mu = __ddiv_rn(sxp, sp);
        double pnd = __ddiv_rn(snd, (double)N);
        fp::rowtree::leaf_sums(StagedExpressedDetect{mu, params, ldp, row}, leaves, n_leaves, val);
        __syncthreads();
        const double mean_pde = __ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N);
        const double pe = __ddiv_rn(__dsub_rn(1.0, pnd), mean_pde);
        if (pnd == 0.0) pnd = __ddiv_rn(1.0, (double)N);
        double *orow = out + g * ld_out;
        for (long c = threadIdx.x; c < N; c += blockDim.x) {
            const double p_nd = (zero_bits[c >> 5] >> (c & 31)) & 1u ? 1.0 : 0.0;
            const double p_d = __dsub_rn(1.0, p_nd);
            double v = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(__dsub_rn(1.0, row[c]), pe), pnd));
            if (v < 0.0) v = 0.0;
            if (v > 1.0) v = 1.0;
            orow[c] = __dadd_rn(__dmul_rn(v, p_nd), p_d);
        }
        __syncthreads();                                  // the row buffer is free again
        if (threadIdx.x == 0 && next < G && n_bulk)
            fp::rowtree::stage_row_async(row, X + next * ld, (uint32_t)(n_bulk * 8), &bar);
    }
}

}  // namespace

extern "C" size_t fp_compute_weights_workspace_bytes(int64_t N) {
    if (N <= 0) return 0;
    return fp::rowtree::workspace_bytes(N);
}

extern "C" int fp_compute_weights(const double *X, const double *p_nd, const double *params, int64_t G, int64_t N,
                                  int64_t ld, int64_t ld_pnd, double *out_weights, int64_t ld_out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(X && params && out_weights, "fp_compute_weights: null pointer");
    FP_CHECK_ARG(G > 0 && N > 0 && ld >= N && ld_out >= N && (!p_nd || ld_pnd >= N),
                 "fp_compute_weights: bad shape G=%ld N=%ld", (long)G, (long)N);
    FP_CHECK_ARG(N < (1L << 31), "fp_compute_weights: N too large");
    std::vector<Leaf> leaves;
    std::vector<int> prog;
    int max_depth = 0;
    build_tree(0, N, leaves, prog, 1, max_depth);
    const size_t smem = 2 * leaves.size() * sizeof(double);
    if (max_depth + 1 > MAX_DEPTH || smem > 200 * 1024) {
        fp::set_error("fp_compute_weights: rows of %ld cells need %zu tree leaves (limit 12800)", (long)N,
                      leaves.size());
        return FP_EUNSUPPORTED;
    }
    const size_t leaves_b = align_up(leaves.size() * sizeof(Leaf)), prog_b = align_up(prog.size() * sizeof(int)),
                 nodes_b = align_up(leaves.size() * sizeof(Node));
    if (!workspace || workspace_bytes < leaves_b + prog_b + nodes_b) {
        fp::set_error("fp_compute_weights: workspace %zu bytes < required %zu", workspace_bytes,
                      leaves_b + prog_b + nodes_b);
        return FP_EWORKSPACE;
    }
    cudaStream_t st = fp::as_stream(stream);
    char *ws = static_cast<char *>(workspace);
    const Leaf *d_leaves = reinterpret_cast<const Leaf *>(ws);
    const Node *d_nodes = reinterpret_cast<const Node *>(ws + leaves_b + prog_b);
    fp::rowtree::build_tree_kernel<<<1, 32, 0, st>>>(N, reinterpret_cast<Leaf *>(ws), reinterpret_cast<int *>(ws + leaves_b),
                                                     reinterpret_cast<Node *>(ws + leaves_b + prog_b));
    FP_CHECK_LAUNCH("build_tree_kernel");              // the host copy above only sized the tables
    const int height = max_depth - 1;
    const size_t staged_b = (size_t)((N + 1) & ~1L) * sizeof(double) + 2 * leaves.size() * sizeof(double) +
                            (size_t)((N + 31) / 32) * sizeof(uint32_t);
    if (!p_nd && fp::rowtree::stage_applies(X, ld, staged_b)) {
        static size_t staged_set = 0;
        if (staged_b > staged_set) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(compute_weights_staged_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_b));
            staged_set = staged_b;
        }
        long blocks = (long)fp::sm_count() * (staged_b <= 100 * 1024 ? 2 : 1);
        if (blocks > G) blocks = G;
        compute_weights_staged_kernel<<<(unsigned)blocks, STAGED_THREADS, staged_b, st>>>(
            X, params, N, G, N, ld, out_weights, ld_out, d_leaves, (int)leaves.size(), d_nodes, height);
        FP_CHECK_LAUNCH("compute_weights_staged_kernel");
        return FP_OK;
    }
    static size_t smem_set = 0;
    if (smem > 40 * 1024 && smem > smem_set) {
        FP_CHECK_CUDA(cudaFuncSetAttribute(compute_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        smem_set = smem;
    }
    long blocks = G;
    const long cap = (long)fp::sm_count() * 4;
    if (blocks > cap) blocks = cap;
    compute_weights_kernel<<<(unsigned)blocks, W_THREADS, smem, st>>>(
        X, p_nd, params, N, G, N, ld, ld_pnd, out_weights, ld_out, d_leaves, (int)leaves.size(), d_nodes, height);
    FP_CHECK_LAUNCH("compute_weights_kernel");
    return FP_OK;
}
