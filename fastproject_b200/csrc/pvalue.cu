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
#include "fp_common.cuh"

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

__global__ void background_stats_kernel(const double *__restrict__ rnd_med,
                                        const int32_t *__restrict__ rnd_order,
                                        const int32_t *__restrict__ group_ptr, long n_groups,
                                        double *__restrict__ out_mu, double *__restrict__ out_sigma) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const long lo = group_ptr[g], n = group_ptr[g + 1] - lo;
    GroupValue val{rnd_med, rnd_order};
    const double mean = __ddiv_rn(fp::np_pairwise_sum(val, lo, n), (double)n);
    GroupSquaredDev dev{rnd_med, rnd_order, mean};
    const double var = __ddiv_rn(fp::np_pairwise_sum(dev, lo, n), (double)n);
    out_mu[g] = mean;
    out_sigma[g] = sqrt(var);
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
        background_stats_kernel<<<(unsigned)((n_groups + 31) / 32), 32, 0, st>>>(
            rnd_med, rnd_order, group_ptr, n_groups, out_mu, out_sigma);
        FP_CHECK_LAUNCH("background_stats_kernel");
    }
    if (K > 0) {
        pvalue_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(med, sig_group, K, out_mu,
                                                                   out_sigma, out_p);
        FP_CHECK_LAUNCH("pvalue_kernel");
    }
    return FP_OK;
}
