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

// Keys of a row: from the shared-memory copy when the row fits (CACHED), else from
// global memory (L2) with four independent loads in flight per thread.
template <bool CACHED>
struct RowKeys {
    const double *row;
    const unsigned long long *cache;
    __device__ __forceinline__ unsigned long long operator()(long i) const {
        return CACHED ? cache[i] : bits_of(row[i]);
    }
};

// k-th smallest (0-based) bit pattern of the row; all threads return it.
template <bool CACHED>
__device__ unsigned long long radix_select(const RowKeys<CACHED> keys, long n, long k,
                                           unsigned *hist /*256*/, long *bcast /*2*/) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        // run-length aggregation: the high bytes of a row are nearly constant, so a
        // thread flushes one atomic per run of equal digits instead of one per element
        unsigned run_digit = 0, run_len = 0;
        const long stride = blockDim.x;
        long i = threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            unsigned long long u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] = keys(i + q * stride);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((u[q] & mask) == prefix) {
                    const unsigned d = (unsigned)(u[q] >> shift) & 255u;
                    if (d != run_digit && run_len) {
                        atomicAdd(&hist[run_digit], run_len);
                        run_len = 0;
                    }
                    run_digit = d;
                    ++run_len;
                }
            }
        }
        for (; i < n; i += stride) {
            const unsigned long long u = keys(i);
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

template <bool CACHED>
__global__ void __launch_bounds__(THREADS)
row_median_kernel(const double *__restrict__ values, long K, long N, long ld,
                  double *__restrict__ out_median) {
    extern __shared__ unsigned long long key_cache[];     // N keys when CACHED
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
        {
            long i = threadIdx.x;
            for (; i + 3 * (long)THREADS < N; i += 4 * THREADS) {
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = row[i + q * THREADS];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    saw_nan |= (v[q] != v[q]);
                    if (CACHED) key_cache[i + q * THREADS] = bits_of(v[q]);
                }
            }
            for (; i < N; i += THREADS) {
                const double v = row[i];
                saw_nan |= (v != v);
                if (CACHED) key_cache[i] = bits_of(v);
            }
        }
        if (saw_nan) nan_flag = 1;
        __syncthreads();
        if (nan_flag) {
            if (threadIdx.x == 0) out_median[row_id] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();
            continue;
        }
        const RowKeys<CACHED> keys{row, key_cache};
        const long k_hi = N / 2;                       // upper middle (the median when N is odd)
        if (N & 1) {
            const unsigned long long u = radix_select<CACHED>(keys, N, k_hi, hist, bcast);
            if (threadIdx.x == 0) out_median[row_id] = value_of(u);
        } else {
            // lower middle = (N/2 - 1)-th; upper middle = it again if enough copies
            // of it exist, else the smallest value above it
            const unsigned long long lo = radix_select<CACHED>(keys, N, k_hi - 1, hist, bcast);
            long cnt_le = 0;
            unsigned long long min_gt = ~0ull;
            for (long i = threadIdx.x; i < N; i += blockDim.x) {
                const unsigned long long u = keys(i);
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
    // a row of up to 25 600 cells is copied once into shared memory (order-preserving keys)
    // and all eight radix passes run from there; longer rows are re-read from L2
    const size_t cache_bytes = (size_t)N * 8;
    // (measured on B200, N = 20 000: the cached variant leaves one CTA per SM and is slower,
    //  1.31 ms vs 1.08 ms per 3 500 rows; kept for short rows only, where several CTAs still fit)
    if (cache_bytes <= 40 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(row_median_kernel<true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        const long per_sm = (long)(220 * 1024 / (cache_bytes + 8192));
        const long cap = (long)fp::sm_count() * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm)) * 8;
        if (blocks > cap) blocks = cap;
        row_median_kernel<true><<<(unsigned)blocks, THREADS, cache_bytes, fp::as_stream(stream)>>>(
            values, K, N, ld, out_median);
    } else {
        const long cap = (long)fp::sm_count() * 4 * 8;
        if (blocks > cap) blocks = cap;
        row_median_kernel<false><<<(unsigned)blocks, THREADS, 0, fp::as_stream(stream)>>>(values, K, N, ld,
                                                                                          out_median);
    }
    FP_CHECK_LAUNCH("row_median_kernel");
    return FP_OK;
}
