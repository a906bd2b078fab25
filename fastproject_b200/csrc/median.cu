// Exact per-signature median of the local dissimilarities.
//
// Replaces np.median(dissimilarity, axis=0) and its random-signature twin
// (/root/reference/FastProject/Signatures.py:409, 414).  A median is a
// selection, not a reduction: one CTA per signature row runs an MSB-first
// radix select (8 passes of 8 bits) over order-preserving keys of the IEEE-754
// bit patterns (sign bit flipped for non-negative values, all bits flipped for
// negative ones), so any doubles are handled, not only absolute values.
// Even N: mean of the two middle order statistics, (a + b) / 2 as NumPy does.
// A NaN anywhere in the row gives NaN (np.median propagates NaN).
#include "fp_common.cuh"

namespace {

constexpr int THREADS = 512;

__device__ __forceinline__ unsigned long long bits_of(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long key) {
    const unsigned long long u = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
    return __longlong_as_double((long long)u);
}

// k-th smallest (0-based) bit pattern of row[0..n); all threads return it.
__device__ unsigned long long radix_select(const double *__restrict__ row, long n, long k,
                                           unsigned *hist /*256*/, long *bcast /*2*/) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        // run-length aggregation: the high bytes of a row are nearly constant, so a
        // thread flushes one atomic per run of equal digits instead of one per element
        unsigned run_digit = 0, run_len = 0;
        for (long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long u = bits_of(row[i]);
            if ((u & mask) == prefix) {
                const unsigned d = (unsigned)(u >> shift) & 255u;
                if (d != run_digit && run_len) {
                    atomicAdd(&hist[run_digit], run_len);
                    run_len = 0;
                }
                run_digit = d;
                ++run_len;
            }
        }
        if (run_len) atomicAdd(&hist[run_digit], run_len);
        __syncthreads();
        if (threadIdx.x == 0) {
            long below = 0;
            int b = 0;
            for (; b < 256; ++b) {
                const long c = hist[b];
                if (below + c > k) break;
                below += c;
            }
            bcast[0] = b;
            bcast[1] = k - below;
        }
        __syncthreads();
        prefix |= (unsigned long long)bcast[0] << shift;
        mask |= 0xffull << shift;
        k = bcast[1];
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(THREADS)
row_median_kernel(const double *__restrict__ values, long K, long N, long ld,
                  double *__restrict__ out_median) {
    __shared__ unsigned hist[256];
    __shared__ long bcast[2];
    __shared__ unsigned long long red_min[THREADS / 32];
    __shared__ long red_cnt[THREADS / 32];
    __shared__ int nan_flag;
    for (long row_id = blockIdx.x; row_id < K; row_id += gridDim.x) {
        const double *row = values + row_id * ld;
        if (threadIdx.x == 0) nan_flag = 0;
        __syncthreads();
        int saw_nan = 0;
        for (long i = threadIdx.x; i < N; i += blockDim.x) {
            const double v = row[i];
            saw_nan |= (v != v);
        }
        if (saw_nan) nan_flag = 1;
        __syncthreads();
        if (nan_flag) {
            if (threadIdx.x == 0) out_median[row_id] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();
            continue;
        }
        const long k_hi = N / 2;                       // upper middle (the median when N is odd)
        if (N & 1) {
            const unsigned long long u = radix_select(row, N, k_hi, hist, bcast);
            if (threadIdx.x == 0) out_median[row_id] = value_of(u);
        } else {
            // lower middle = (N/2 - 1)-th; upper middle = it again if enough copies
            // of it exist, else the smallest value above it
            const unsigned long long lo = radix_select(row, N, k_hi - 1, hist, bcast);
            long cnt_le = 0;
            unsigned long long min_gt = ~0ull;
            for (long i = threadIdx.x; i < N; i += blockDim.x) {
                const unsigned long long u = bits_of(row[i]);
                if (u <= lo) ++cnt_le;
                else if (u < min_gt) min_gt = u;
            }
            for (int off = 16; off; off >>= 1) {
                cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, off);
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, min_gt, off);
                if (o < min_gt) min_gt = o;
            }
            if ((threadIdx.x & 31) == 0) {
                red_cnt[threadIdx.x >> 5] = cnt_le;
                red_min[threadIdx.x >> 5] = min_gt;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                long c = 0;
                unsigned long long m = ~0ull;
                for (int w = 0; w < THREADS / 32; ++w) {
                    c += red_cnt[w];
                    if (red_min[w] < m) m = red_min[w];
                }
                const unsigned long long hi = (c > k_hi) ? lo : m;
                const double a = value_of(lo);
                const double b = value_of(hi);
                out_median[row_id] = __dmul_rn(__dadd_rn(a, b), 0.5);
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int fp_row_medians(const double *values, int64_t K, int64_t N, int64_t ld,
                              double *out_median, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(values && out_median, "fp_row_medians: null pointer");
    FP_CHECK_ARG(K >= 0 && N > 0 && ld >= N, "fp_row_medians: bad shape K=%ld N=%ld ld=%ld", (long)K,
                 (long)N, (long)ld);
    if (K == 0) return FP_OK;
    long blocks = K;
    const long cap = (long)fp::sm_count() * 4 * 8;
    if (blocks > cap) blocks = cap;
    row_median_kernel<<<(unsigned)blocks, THREADS, 0, fp::as_stream(stream)>>>(values, K, N, ld,
                                                                               out_median);
    FP_CHECK_LAUNCH("row_median_kernel");
    return FP_OK;
}
