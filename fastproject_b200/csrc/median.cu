// Exact per-signature median of the local dissimilarities.
// (short rows: byte-wise radix select below; long rows: row_median_wide_kernel)
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

// ---------------------------------------------------------------------------
// Long rows: two wide radix passes (12 + 11 bits: sign, exponent and the top mantissa
// bits) narrow the row to the few elements that share the median's top 23 bits; a third
// pass collects them into shared memory, where the remaining 41 bits are resolved.  Three
// sweeps over the row instead of nine (NaN scan + 8 digit passes) or ten (even N).
// Falls back to the byte-wise select when more than CAND_MAX elements share those bits.
// ---------------------------------------------------------------------------
constexpr int CAND_MAX = 3072;
constexpr int WIDE_BINS = 4096;

template <int BITS>
__device__ void wide_histogram(const double *__restrict__ row, long n, int shift, unsigned long long mask,
                               unsigned long long prefix, unsigned *hist, int *nan_flag) {
    for (int b = threadIdx.x; b < (1 << BITS); b += blockDim.x) hist[b] = 0;
    __syncthreads();
    // sign + exponent digits take only a handful of values, all hot: a four-entry counting
    // cache in registers absorbs them (an atomic per element on two or three shared-memory
    // words serialises the block); whatever is evicted goes to the histogram
    unsigned cd[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, cc[4] = {0, 0, 0, 0};
    int saw_nan = 0;
    const long stride = blockDim.x;
    auto visit = [&](double v) {
        saw_nan |= (v != v);
        const unsigned long long u = bits_of(v);
        if ((u & mask) == prefix) {
            const unsigned d = (unsigned)(u >> shift) & ((1u << BITS) - 1u);
            if (d == cd[0]) ++cc[0];
            else if (d == cd[1]) ++cc[1];
            else if (d == cd[2]) ++cc[2];
            else if (d == cd[3]) ++cc[3];
            else {
                // evict the oldest entry (slot 3), shift, insert in front
                if (cc[3]) atomicAdd(&hist[cd[3]], cc[3]);
                cd[3] = cd[2]; cc[3] = cc[2];
                cd[2] = cd[1]; cc[2] = cc[1];
                cd[1] = cd[0]; cc[1] = cc[0];
                cd[0] = d; cc[0] = 1;
            }
        }
    };
    long i = threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = row[i + q * stride];
#pragma unroll
        for (int q = 0; q < 4; ++q) visit(v[q]);
    }
    for (; i < n; i += stride) visit(row[i]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (cc[q]) atomicAdd(&hist[cd[q]], cc[q]);
    if (saw_nan) *nan_flag = 1;
    __syncthreads();
}

// bin holding rank k (0-based) and the number of elements below that bin: block-wide prefix
// sum over the histogram (a serial scan of 4096 bins by one thread cost more than the sweeps)
template <int BINS>
__device__ void locate_bin(const unsigned *hist, long k, long *bcast, unsigned *warp_tot /*THREADS/32*/) {
    constexpr int PER = BINS / THREADS;
    static_assert(BINS % THREADS == 0, "bins per thread");
    unsigned local[PER];
    unsigned sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        local[q] = hist[threadIdx.x * PER + q];
        sum += local[q];
    }
    unsigned incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    long below = 0;
    for (int w = 0; w < warp; ++w) below += warp_tot[w];
    below += incl - sum;                               // exclusive prefix of this thread's bins
    if (below <= k && k < below + (long)sum) {
        int q = 0;
        for (; q < PER; ++q) {
            if (below + (long)local[q] > k) break;
            below += local[q];
        }
        bcast[0] = threadIdx.x * PER + q;
        bcast[1] = below;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS)
row_median_wide_kernel(const double *__restrict__ values, long K, long N, long ld,
                       double *__restrict__ out_median) {
    __shared__ unsigned hist[WIDE_BINS];
    __shared__ unsigned long long cand[CAND_MAX];
    __shared__ long bcast[2];
    __shared__ unsigned long long red_min[THREADS / 32];
    __shared__ int nan_flag;
    __shared__ unsigned n_cand;
    __shared__ unsigned warp_tot[THREADS / 32];
    const long k_hi = N / 2, k_lo = (N & 1) ? k_hi : k_hi - 1;
    for (long row_id = blockIdx.x; row_id < K; row_id += gridDim.x) {
        const double *row = values + row_id * ld;
        if (threadIdx.x == 0) {
            nan_flag = 0;
            n_cand = 0;
        }
        __syncthreads();
        // pass A: top 12 bits (sign + exponent); also the NaN scan
        wide_histogram<12>(row, N, 52, 0ull, 0ull, hist, &nan_flag);
        if (nan_flag) {
            if (threadIdx.x == 0) out_median[row_id] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();
            continue;
        }
        locate_bin<4096>(hist, k_lo, bcast, warp_tot);
        unsigned long long prefix = (unsigned long long)bcast[0] << 52;
        long below = bcast[1];
        __syncthreads();
        // pass B: next 11 bits among the elements of that bin
        wide_histogram<11>(row, N, 41, 0xfffull << 52, prefix, hist, &nan_flag);
        locate_bin<2048>(hist, k_lo - below, bcast, warp_tot);
        prefix |= (unsigned long long)bcast[0] << 41;
        below += bcast[1];
        const long in_bin = hist[bcast[0]];
        __syncthreads();
        if (in_bin > CAND_MAX) {
            // many elements share the median's top 23 bits (e.g. a constant row): byte-wise select
            const RowKeys<false> keys{row, nullptr};
            const unsigned long long lo = radix_select<false>(keys, N, k_lo, hist, bcast);
            unsigned long long hi = lo;
            if (k_hi != k_lo) {
                long cnt_le = 0;
                unsigned long long min_gt = ~0ull;
                for (long i = threadIdx.x; i < N; i += blockDim.x) {
                    const unsigned long long u = keys(i);
                    if (u <= lo) ++cnt_le;
                    else if (u < min_gt) min_gt = u;
                }
                // block reduction through shared memory (hist is free now)
                unsigned long long *cnts = reinterpret_cast<unsigned long long *>(hist);
                for (int off = 16; off; off >>= 1) {
                    cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, off);
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, min_gt, off);
                    if (o < min_gt) min_gt = o;
                }
                __syncthreads();
                if ((threadIdx.x & 31) == 0) {
                    cnts[threadIdx.x >> 5] = (unsigned long long)cnt_le;
                    red_min[threadIdx.x >> 5] = min_gt;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    unsigned long long c = 0, m = ~0ull;
                    for (int w = 0; w < THREADS / 32; ++w) {
                        c += cnts[w];
                        if (red_min[w] < m) m = red_min[w];
                    }
                    bcast[0] = (long)(((long)c > k_hi) ? lo : m);
                }
                __syncthreads();
                hi = (unsigned long long)bcast[0];
            }
            if (threadIdx.x == 0) out_median[row_id] = __dmul_rn(__dadd_rn(value_of(lo), value_of(hi)), 0.5);
            __syncthreads();
            continue;
        }
        // pass C: collect the bin; remember the smallest key above it (upper middle of an even N
        // when the lower middle is the bin's largest element)
        unsigned long long min_above = ~0ull;
        {
            const unsigned long long mask23 = ~((1ull << 41) - 1ull);
            long i = threadIdx.x;
            const long stride = blockDim.x;
            auto visit = [&](double v) {
                const unsigned long long u = bits_of(v);
                const unsigned long long top = u & mask23;
                if (top == prefix) cand[atomicAdd(&n_cand, 1u)] = u;
                else if (top > prefix && u < min_above) min_above = u;
            };
            for (; i + 3 * stride < N; i += 4 * stride) {
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = row[i + q * stride];
#pragma unroll
                for (int q = 0; q < 4; ++q) visit(v[q]);
            }
            for (; i < N; i += stride) visit(row[i]);
        }
        for (int off = 16; off; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, min_above, off);
            if (o < min_above) min_above = o;
        }
        if ((threadIdx.x & 31) == 0) red_min[threadIdx.x >> 5] = min_above;
        __syncthreads();
        const long nc = n_cand;
        unsigned long long lo, hi;
        if (nc <= 2 * THREADS) {
            // a handful of candidates: every thread ranks its own by counting (ties by position)
            __shared__ unsigned long long picked[2];
            for (long idx = threadIdx.x; idx < nc; idx += blockDim.x) {
                const unsigned long long mine = cand[idx];
                long r = 0;
                for (long j = 0; j < nc; ++j) {
                    const unsigned long long o = cand[j];
                    r += (o < mine) || (o == mine && j < idx);
                }
                if (r == k_lo - below) picked[0] = mine;
                if (r == k_hi - below) picked[1] = mine;
            }
            __syncthreads();
            lo = picked[0];
            if (k_hi - below < nc) {
                hi = picked[1];
            } else {
                unsigned long long m = ~0ull;
                for (int w = 0; w < THREADS / 32; ++w)
                    if (red_min[w] < m) m = red_min[w];
                hi = m;
            }
        } else {
            const RowKeys<true> ckeys{nullptr, cand};
            lo = radix_select<true>(ckeys, nc, k_lo - below, hist, bcast);
            hi = lo;
            if (k_hi != k_lo) {
                if (k_hi - below < nc) {
                    hi = radix_select<true>(ckeys, nc, k_hi - below, hist, bcast);
                } else {
                    unsigned long long m = ~0ull;
                    for (int w = 0; w < THREADS / 32; ++w)
                        if (red_min[w] < m) m = red_min[w];
                    hi = m;
                }
            }
        }
        if (threadIdx.x == 0) out_median[row_id] = __dmul_rn(__dadd_rn(value_of(lo), value_of(hi)), 0.5);
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
    // short rows (up to 2 048 cells) are copied once into shared memory and all eight byte-wise
    // radix passes run from there; longer rows take the wide-digit kernel (three sweeps from L2).
    // Measured on B200, 3 500 rows of 20 000 cells: byte-wise from L2 1.08 ms (0.87 with unrolled
    // loads), byte-wise from a shared-memory copy 1.31 ms (one CTA per SM), wide digits 0.47 ms.
    const size_t cache_bytes = (size_t)N * 8;
    if (cache_bytes <= 16 * 1024) {
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
        row_median_wide_kernel<<<(unsigned)blocks, THREADS, 0, fp::as_stream(stream)>>>(values, K, N, ld,
                                                                                        out_median);
    }
    FP_CHECK_LAUNCH("row_median_kernel");
    return FP_OK;
}
