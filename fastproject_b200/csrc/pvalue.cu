// Random-background statistics and p-values of the real signatures.
//
// Replaces /root/reference/FastProject/Signatures.py:417-442 for one
// projection: per-numGenes groups of the random signatures' median
// dissimilarities -> (mu, sigma) with np.mean / np.std semantics (population
// std, NumPy pairwise summation order, so mu and sigma are bit-identical to
// NumPy's given identical medians), then p = Phi((med - mu) / sigma) following
// scipy.special.ndtr's branch structure (erf inside |x/sqrt2| < 1, erfc
// outside).  sigma == 0 is NOT guarded, exactly as in the reference (:442):
// the quotient becomes +-inf or NaN and the CDF follows IEEE rules.
#include "row_tree.cuh"

namespace {

struct GroupValue {
    const double *rnd_med;
    const int32_t *order;
    __device__ double operator()(long i) const { return rnd_med[order[i]]; }
};
struct GroupSquaredDev {
    const double *rnd_med;
    const int32_t *order;
    double mean;
    __device__ double operator()(long i) const {
        // np.std: abs(x - mean)**2 -- for reals NumPy multiplies the deviation by itself
        const double d = __dsub_rn(rnd_med[order[i]], mean);
        return __dmul_rn(d, d);
    }
};

// Block form: one CTA per group.  Thread 0 enumerates the leaves of NumPy's pairwise tree for
// this group's size and its post-order program (the recursion of fp::np_pairwise_sum, made
// iterative), eight lanes sum each leaf, thread 0 folds.  A group of 500 medians took 116 us
// with one thread walking both sums.
constexpr int STATS_THREADS = 256;
constexpr int STATS_MAX_LEAVES = 1024;

using StatLeaf = fp::rowtree::Leaf;

template <typename Load>
__device__ double stats_tree_sum(Load load, const StatLeaf *leaves, int n_leaves, const int *prog, int n_prog,
                                 double *leaf_sum, double *bcast) {
    if (n_leaves == 1 && leaves[0].n < 8) {
        if (threadIdx.x == 0) {                  // short loop, from -0.0 like NumPy's
            double res = -0.0;
            for (int i = 0; i < leaves[0].n; ++i) res = __dadd_rn(res, load(leaves[0].lo + i));
            leaf_sum[0] = res;
        }
    } else {
        const int per_pass = STATS_THREADS >> 3;
        for (int base = 0; base < n_leaves; base += per_pass) {
            const int l = base + (threadIdx.x >> 3);
            const bool live = l < n_leaves;
            const StatLeaf leaf = leaves[live ? l : 0];
            const double v = fp::np_leaf_sum_octet(load, leaf.lo, leaf.n);
            if (live && (threadIdx.x & 7) == 0) leaf_sum[l] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double stack[48];
        int sp = 0;
        for (int p = 0; p < n_prog; ++p) {
            const int op = prog[p];
            if (op >= 0) stack[sp++] = leaf_sum[op];
            else {
                const double b = stack[--sp], a = stack[--sp];
                stack[sp++] = __dadd_rn(a, b);
            }
        }
        *bcast = stack[0];
    }
    __syncthreads();
    return *bcast;
}

__global__ void __launch_bounds__(STATS_THREADS)
background_stats_kernel(const double *__restrict__ rnd_med, const int32_t *__restrict__ rnd_order,
                        const int32_t *__restrict__ group_ptr, long n_groups, double *__restrict__ out_mu,
                        double *__restrict__ out_sigma) {
    __shared__ StatLeaf leaves[STATS_MAX_LEAVES];
    __shared__ int prog[2 * STATS_MAX_LEAVES];
    __shared__ double leaf_sum[STATS_MAX_LEAVES];
    __shared__ int n_leaves, n_prog;
    __shared__ double bcast;
    const long g = blockIdx.x;
    const long lo = group_ptr[g], n = group_ptr[g + 1] - lo;
    GroupValue val{rnd_med, rnd_order};
    if (n > 60L * STATS_MAX_LEAVES) {
        // more leaves than the table holds (a leaf has at least 60 elements): one thread, serially
        if (threadIdx.x == 0) {
            const double mean = __ddiv_rn(fp::np_pairwise_sum(val, lo, n), (double)n);
            GroupSquaredDev dev{rnd_med, rnd_order, mean};
            out_mu[g] = mean;
            out_sigma[g] = sqrt(__ddiv_rn(fp::np_pairwise_sum(dev, lo, n), (double)n));
        }
        return;
    }
    if (threadIdx.x == 0) fp::rowtree::build_tree_device(lo, n, leaves, &n_leaves, prog, &n_prog);
    __syncthreads();
    const double mean = __ddiv_rn(stats_tree_sum(val, leaves, n_leaves, prog, n_prog, leaf_sum, &bcast), (double)n);
    GroupSquaredDev dev{rnd_med, rnd_order, mean};
    const double var = __ddiv_rn(stats_tree_sum(dev, leaves, n_leaves, prog, n_prog, leaf_sum, &bcast), (double)n);
    if (threadIdx.x == 0) {
        out_mu[g] = mean;
        out_sigma[g] = sqrt(var);
    }
}

__device__ double normal_cdf(double z) {
    const double x = z * 0.70710678118654752440;   // M_SQRT1_2
    if (x != x) return x;
    const double ax = fabs(x);
    if (ax < 1.0) return 0.5 + 0.5 * erf(x);
    const double y = 0.5 * erfc(ax);
    return x > 0 ? 1.0 - y : y;
}

__global__ void pvalue_kernel(const double *__restrict__ med, const int32_t *__restrict__ sig_group,
                              long K, const double *__restrict__ mu, const double *__restrict__ sigma,
                              double *__restrict__ out_p) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int g = sig_group[k];
    out_p[k] = normal_cdf(__ddiv_rn(__dsub_rn(med[k], mu[g]), sigma[g]));
}

}  // namespace

extern "C" int fp_background_pvalues(const double *med, const int32_t *sig_group, int64_t K,
                                     const double *rnd_med, const int32_t *rnd_order,
                                     const int32_t *group_ptr, int64_t n_groups, double *out_p,
                                     double *out_mu, double *out_sigma, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(K >= 0 && n_groups >= 0, "fp_background_pvalues: negative size");
    FP_CHECK_ARG(n_groups == 0 || (rnd_med && rnd_order && group_ptr && out_mu && out_sigma),
                 "fp_background_pvalues: null background pointer");
    FP_CHECK_ARG(K == 0 || (med && sig_group && out_p), "fp_background_pvalues: null pointer");
    FP_CHECK_ARG(K == 0 || n_groups > 0,
                 "fp_background_pvalues: no background groups (the reference fails here too: "
                 "argmin of an empty sequence, Signatures.py:437)");
    cudaStream_t st = fp::as_stream(stream);
    if (n_groups > 0) {
        background_stats_kernel<<<(unsigned)n_groups, STATS_THREADS, 0, st>>>(rnd_med, rnd_order, group_ptr, n_groups,
                                                                             out_mu, out_sigma);
        FP_CHECK_LAUNCH("background_stats_kernel");
    }
    if (K > 0) {
        pvalue_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(med, sig_group, K, out_mu,
                                                                   out_sigma, out_p);
        FP_CHECK_LAUNCH("pvalue_kernel");
    }
    return FP_OK;
}
