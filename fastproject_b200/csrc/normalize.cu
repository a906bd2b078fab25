// Data normalisation before signature scoring -- the step right before the hot path
// (SURVEY.md 8f, row N1).  Replaces /root/reference/FastProject/NormalizationMethods.py:
//   row_normalization (:34-41, the CLI default "znorm_rows"), col_normalization (:24-31);
// row_and_col_normalization (:44-50) is the two applied in turn by the Python mirror and
// col_rank_normalization (:53-61) goes through fp_rank_columns_ex (method "min").
//
// Bit-exactness contract.  NumPy computes
//   mu    = add.reduce(x, axis) / n
//   sigma = sqrt(add.reduce((x - mu) * (x - mu), axis) / n),   sigma == 0 -> 1
//   out   = (x - mu) / sigma
// where add.reduce along the CONTIGUOUS axis (rows, axis=1) is NumPy's pairwise summation
// (blocks of <= 128 elements with eight interleaved accumulators, halves combined
// recursively) and along axis=0 it is a plain sequential sum over the rows.  Both orders are
// reproduced, every operation separately rounded, so the result equals the reference's bit
// for bit.
//
// Row kernel: one CTA per gene.  The leaves of the pairwise tree (it depends on N only; the
// host enumerates it once per call) are summed by one thread each -- a leaf is 1 KB of
// consecutive doubles, so a warp streams 32 leaves through L1 -- and thread 0 folds the leaf
// sums with the tree's post-order program.
#include "row_tree.cuh"

namespace {

using fp::rowtree::Leaf;
using fp::rowtree::Node;
using fp::rowtree::fold_levels;
using fp::rowtree::build_tree;
using fp::rowtree::MAX_DEPTH;
using fp::rowtree::align_up;

struct RowLoad {
    const double *x;
    __device__ double operator()(long i) const { return x[i]; }
};
struct RowSqDev {
    const double *x;
    double mean;
    __device__ double operator()(long i) const {
        const double d = __dsub_rn(x[i], mean);
        return __dmul_rn(d, d);
    }
};

constexpr int ROW_THREADS = 256;

__global__ void __launch_bounds__(ROW_THREADS)
row_zscore_kernel(const double *__restrict__ X, long G, long N, long ld, double *__restrict__ out, long ld_out,
                  const Leaf *__restrict__ leaves, int n_leaves, const Node *__restrict__ nodes, int height) {
    extern __shared__ double val[];                      // 2 * n_leaves
    for (long g = blockIdx.x; g < G; g += gridDim.x) {
        const double *row = X + g * ld;
        fp::rowtree::leaf_sums(RowLoad{row}, leaves, n_leaves, val);
        __syncthreads();
        const double mean = __ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N);
        fp::rowtree::leaf_sums(RowSqDev{row, mean}, leaves, n_leaves, val);
        __syncthreads();
        double sd = sqrt(__ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N));
        if (sd == 0.0) sd = 1.0;                          // data_sigma[data_sigma == 0] = 1.0
        double *orow = out + g * ld_out;
        for (long c = threadIdx.x; c < N; c += blockDim.x) orow[c] = __ddiv_rn(__dsub_rn(row[c], mean), sd);
        __syncthreads();
    }
}

// Rows that fit in shared memory (N <= ~28 000 cells): the row is staged once by bulk
// asynchronous copies (row_tree.cuh), both tree sums and the final sweep read it from there, and
// the next row of the CTA is prefetched into L2 meanwhile -- DRAM sees one read and one write of
// the matrix (the kernel above re-reads each row twice, and with enough rows in flight to cover
// the latency they do not all stay in L2: 6.1 GB read for a 2.4 GB matrix).
constexpr int STAGED_THREADS = 1024;

__global__ void __launch_bounds__(STAGED_THREADS)
row_zscore_staged_kernel(const double *__restrict__ X, long G, long N, long ld, double *__restrict__ out, long ld_out,
                         const Leaf *__restrict__ leaves, int n_leaves, const Node *__restrict__ nodes, int height) {
    extern __shared__ __align__(16) double sm[];         // row (N rounded up to even) | 2 * n_leaves
    __shared__ uint64_t bar;
    double *row = sm, *val = sm + ((N + 1) & ~1L);
    const long n_bulk = N & ~1L;                         // bulk copies move multiples of 16 bytes
    if (threadIdx.x == 0) {
        fp::rowtree::stage_init(&bar);
        if (n_bulk && blockIdx.x < G) fp::rowtree::stage_row_async(row, X + blockIdx.x * ld, (uint32_t)(n_bulk * 8), &bar);
    }
    __syncthreads();
    uint32_t parity = 0;
    for (long g = blockIdx.x; g < G; g += gridDim.x) {
        const double *src = X + g * ld;
        const long next = g + gridDim.x;
        if (threadIdx.x == 32 && next < G && n_bulk) fp::rowtree::prefetch_row_l2(X + next * ld, (uint32_t)(n_bulk * 8));
        if (n_bulk) {
            fp::rowtree::stage_wait(&bar, parity);
            parity ^= 1;
        }
        if (N & 1) {
            if (threadIdx.x == 0) row[N - 1] = src[N - 1];
            __syncthreads();
        }
        fp::rowtree::leaf_sums(RowLoad{row}, leaves, n_leaves, val);
        __syncthreads();
        const double mean = __ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N);
        fp::rowtree::leaf_sums(RowSqDev{row, mean}, leaves, n_leaves, val);
        __syncthreads();
        double sd = sqrt(__ddiv_rn(fold_levels(val, nodes, n_leaves, height), (double)N));
        if (sd == 0.0) sd = 1.0;
        double *orow = out + g * ld_out;
        for (long c = threadIdx.x; c < N; c += blockDim.x) orow[c] = __ddiv_rn(__dsub_rn(row[c], mean), sd);
        __syncthreads();                                  // the row buffer is free again
        if (threadIdx.x == 0 && next < G && n_bulk)
            fp::rowtree::stage_row_async(row, X + next * ld, (uint32_t)(n_bulk * 8), &bar);
    }
}

// axis = 0: plain sequential sums over the rows, one thread per column (coalesced).  Only
// N threads exist, so every thread keeps COL_DEPTH loads in flight; the additions stay in
// row order.
constexpr int COL_DEPTH = 16;

__global__ void __launch_bounds__(64)
col_zscore_kernel(const double *__restrict__ X, long G, long N, long ld, double *__restrict__ out, long ld_out) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const double *col = X + c;
    double s = 0.0;
    long g = 0;
    for (; g + COL_DEPTH <= G; g += COL_DEPTH) {
        double v[COL_DEPTH];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) v[q] = col[(g + q) * ld];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) s = __dadd_rn(s, v[q]);
    }
    for (; g < G; ++g) s = __dadd_rn(s, col[g * ld]);
    const double mean = __ddiv_rn(s, (double)G);
    double s2 = 0.0;
    for (g = 0; g + COL_DEPTH <= G; g += COL_DEPTH) {
        double v[COL_DEPTH];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) v[q] = col[(g + q) * ld];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) {
            const double d = __dsub_rn(v[q], mean);
            s2 = __dadd_rn(s2, __dmul_rn(d, d));
        }
    }
    for (; g < G; ++g) {
        const double d = __dsub_rn(col[g * ld], mean);
        s2 = __dadd_rn(s2, __dmul_rn(d, d));
    }
    double sd = sqrt(__ddiv_rn(s2, (double)G));
    if (sd == 0.0) sd = 1.0;
    double *ocol = out + c;
    for (g = 0; g + COL_DEPTH <= G; g += COL_DEPTH) {
        double v[COL_DEPTH];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) v[q] = col[(g + q) * ld];
#pragma unroll
        for (int q = 0; q < COL_DEPTH; ++q) ocol[(g + q) * ld_out] = __ddiv_rn(__dsub_rn(v[q], mean), sd);
    }
    for (; g < G; ++g) ocol[g * ld_out] = __ddiv_rn(__dsub_rn(col[g * ld], mean), sd);
}

}  // namespace

extern "C" size_t fp_normalize_rows_workspace_bytes(int64_t N) {
    if (N <= 0) return 0;
    return fp::rowtree::workspace_bytes(N);
}

extern "C" int fp_normalize_rows(const double *X, int64_t G, int64_t N, int64_t ld, double *out, int64_t ld_out,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(X && out, "fp_normalize_rows: null pointer");
    FP_CHECK_ARG(G > 0 && N > 0 && ld >= N && ld_out >= N, "fp_normalize_rows: bad shape G=%ld N=%ld", (long)G,
                 (long)N);
    FP_CHECK_ARG(N < (1L << 31), "fp_normalize_rows: N too large");
    std::vector<Leaf> leaves;
    std::vector<int> prog;
    int max_depth = 0;
    build_tree(0, N, leaves, prog, 1, max_depth);
    if (max_depth + 1 > MAX_DEPTH || leaves.size() * 16 > 200 * 1024) {
        fp::set_error("fp_normalize_rows: rows of %ld cells need %zu tree leaves (limit 12800)", (long)N,
                      leaves.size());
        return FP_EUNSUPPORTED;
    }
    const size_t leaves_b = align_up(leaves.size() * sizeof(Leaf)), prog_b = align_up(prog.size() * sizeof(int)),
                 nodes_b = align_up(leaves.size() * sizeof(Node));
    if (!workspace || workspace_bytes < leaves_b + prog_b + nodes_b) {
        fp::set_error("fp_normalize_rows: workspace %zu bytes < required %zu", workspace_bytes,
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
    const size_t val_b = 2 * leaves.size() * sizeof(double);
    const size_t staged_b = (size_t)((N + 1) & ~1L) * sizeof(double) + val_b;
    if (fp::rowtree::stage_applies(X, ld, staged_b)) {
        static size_t staged_set = 0;
        if (staged_b > staged_set) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(row_zscore_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)staged_b));
            staged_set = staged_b;
        }
        long blocks = (long)fp::sm_count() * (staged_b <= 100 * 1024 ? 2 : 1);
        if (blocks > G) blocks = G;
        row_zscore_staged_kernel<<<(unsigned)blocks, STAGED_THREADS, staged_b, st>>>(X, G, N, ld, out, ld_out, d_leaves,
                                                                                      (int)leaves.size(), d_nodes, height);
        FP_CHECK_LAUNCH("row_zscore_staged_kernel");
        return FP_OK;
    }
    static size_t smem_set = 0;
    if (val_b > 40 * 1024 && val_b > smem_set) {
        FP_CHECK_CUDA(cudaFuncSetAttribute(row_zscore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)val_b));
        smem_set = val_b;
    }
    long blocks = G;
    const long cap = (long)fp::sm_count() * 4;         // rows in flight stay within L2 for the re-reads
    if (blocks > cap) blocks = cap;
    row_zscore_kernel<<<(unsigned)blocks, ROW_THREADS, val_b, st>>>(X, G, N, ld, out, ld_out, d_leaves,
                                                                    (int)leaves.size(), d_nodes, height);
    FP_CHECK_LAUNCH("row_zscore_kernel");
    return FP_OK;
}

extern "C" int fp_normalize_cols(const double *X, int64_t G, int64_t N, int64_t ld, double *out, int64_t ld_out,
                                 void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(X && out, "fp_normalize_cols: null pointer");
    FP_CHECK_ARG(G > 0 && N > 0 && ld >= N && ld_out >= N, "fp_normalize_cols: bad shape G=%ld N=%ld", (long)G,
                 (long)N);
    col_zscore_kernel<<<(unsigned)((N + 63) / 64), 64, 0, fp::as_stream(stream)>>>(X, G, N, ld, out, ld_out);
    FP_CHECK_LAUNCH("col_zscore_kernel");
    return FP_OK;
}
