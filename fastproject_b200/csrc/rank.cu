// Average ranks of every signature's scores over the cells.
//
// Replaces SignatureScores.ranks = scipy.stats.rankdata(scores, "average")
// (/root/reference/FastProject/SigScoreMethods.py:170-178) for all K signatures
// at once, and the rank-matrix assembly of Signatures.py:359-369.
//
// A full sort of the 64-bit scores is not needed to rank them: a monotone
// SHORT key puts every cell within a few positions of its final place, and the
// few cells that share a short key are ordered by looking at their doubles.
//
// Step 1 (quantise_rows_kernel): per row, min / max / NaN scan, then the key
//         q = trunc((v - min) * (2^bits - 1) / (max - min)), bits ~ log2(N) + 3
//         rounded up to whole radix digits.  Every operation is monotone in v,
//         so v < u implies q(v) <= q(u) whatever the rounding.
// Step 2: segmented LSD radix sort of (q, cell) pairs over `bits` bits only
//         (CUB device primitive -- a library call; 3 passes over 8-byte pairs
//         for 20 000 cells instead of 11 passes over 12-byte pairs for the
//         doubles).
// Step 3 (assign_ranks_kernel): a cell alone in its key bin has rank
//         position + 1.  Cells sharing a bin [a, b) count the bin's doubles
//         below and equal to their own: rank = a + less + (equal + 1) / 2
//         ("average") or a + less + 1 ("min").  Comparisons are on double
//         values, so -0.0 ties with +0.0 exactly as in rankdata.
// Step 4 (long_bins_*_kernel): bins longer than kShortBin cells (a constant
//         row, the zeros of an expression column, data squeezed by an outlier)
//         go to a device work list, one warp each: an all-equal bin is ranked
//         in O(L); a mixed one by counting against the whole bin (large mixed
//         bins by a whole CTA).  No host round trip decides between the paths.
// A row holding a NaN yields NaN everywhere (scipy nan_policy="propagate").
// Rows with infinities fall back to keys cut from the order-preserving bit
// pattern of the double (coarse, hence slower through step 4, but exact).
#include <cub/device/device_segmented_radix_sort.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <stdlib.h>
#include <string.h>

#include "fp_common.cuh"

namespace {

constexpr long kMaxBatchElems = 1L << 27;  // elements sorted per CUB call
constexpr int kShortBin = 32;              // bins up to this size are counted in place
constexpr int kWarpBin = 1024;             // mixed bins up to this size are counted by one warp
constexpr int kRadixBits = 6;              // digit width of CUB's segmented sort passes

struct RowBegin {
    long ld;
    __host__ __device__ long operator()(long k) const { return k * ld; }
};
struct RowEnd {
    long ld, n;
    __host__ __device__ long operator()(long k) const { return k * ld + n; }
};

using CountIt = thrust::counting_iterator<long>;
using BeginIt = thrust::transform_iterator<RowBegin, CountIt, long>;
using EndIt = thrust::transform_iterator<RowEnd, CountIt, long>;

struct LongBin {
    int row, a, b, pad;
};

// order-preserving bit pattern of a double (-0.0 folded onto +0.0 first)
__device__ __forceinline__ uint64_t ordered_bits(double v) {
    const uint64_t u = (uint64_t)__double_as_longlong(v + 0.0);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}

constexpr int QUANT_THREADS = 512;

// one CTA per row: range scan, then keys (second read of the row comes from L2)
__global__ void __launch_bounds__(QUANT_THREADS)
quantise_rows_kernel(const double *__restrict__ scores, long rows, long N, long ld, int bits,
                     uint32_t *__restrict__ keys, int32_t *__restrict__ cells, int32_t *__restrict__ row_nan) {
    __shared__ double s_min[QUANT_THREADS / 32], s_max[QUANT_THREADS / 32];
    __shared__ int s_flag[QUANT_THREADS / 32];
    __shared__ double s_lo, s_scale;
    __shared__ int s_mode;                                   // 0 linear keys, 1 bit-pattern keys, 2 NaN row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double top = (double)((1ULL << bits) - 1);
    for (long row = blockIdx.x; row < rows; row += gridDim.x) {
        const double *src = scores + row * ld;
        double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
        int flag = 0;                                        // bit 0: NaN seen, bit 1: infinity seen
        for (long c = threadIdx.x; c < N; c += QUANT_THREADS) {
            const double v = src[c];
            if (v != v) flag |= 1;
            else {
                if (isinf(v)) flag |= 2;
                mn = fmin(mn, v);
                mx = fmax(mx, v);
            }
        }
        for (int o = 16; o; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            flag |= __shfl_xor_sync(0xffffffffu, flag, o);
        }
        if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; s_flag[warp] = flag; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < QUANT_THREADS / 32; ++w) {
                mn = fmin(mn, s_min[w]);
                mx = fmax(mx, s_max[w]);
                flag |= s_flag[w];
            }
            const double span = mx - mn;                     // may overflow to +inf: scale 0, one bin
            s_lo = mn;
            const double scale = (span > 0.0 && !isinf(span)) ? top / span : 0.0;
            s_scale = scale;                                 // a denormal span overflows it: bit-pattern keys
            s_mode = (flag & 1) ? 2 : ((flag & 2) || isinf(span) || isinf(scale)) ? 1 : 0;
            row_nan[row] = (flag & 1);
        }
        __syncthreads();
        const double lo = s_lo, scale = s_scale;
        const int mode = s_mode;
        uint32_t *kd = keys + row * ld;
        int32_t *cd = cells + row * ld;
        for (long c = threadIdx.x; c < ld; c += QUANT_THREADS) {
            uint32_t q = 0;
            if (c < N && mode != 2) {
                const double v = src[c];
                if (mode == 0) {
                    const double t = fmin((v - lo) * scale, top);
                    q = (uint32_t)__double2ull_rz(t);
                } else {
                    q = (uint32_t)(ordered_bits(v) >> (64 - bits));
                }
            }
            kd[c] = q;
            cd[c] = (int32_t)(c < N ? c : 0);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
assign_ranks_kernel(const double *__restrict__ scores, const uint32_t *__restrict__ sorted_keys,
                    const int32_t *__restrict__ sorted_cell, const int32_t *__restrict__ row_nan, long rows, long N,
                    long ld, double *__restrict__ out_ranks, long ld_out, int method, LongBin *__restrict__ work,
                    unsigned *__restrict__ work_count) {
  for (long row = blockIdx.y; row < rows; row += gridDim.y) {
    const uint32_t *ks = sorted_keys + row * ld;
    const int32_t *cs = sorted_cell + row * ld;
    const double *src = scores + row * ld;
    double *out = out_ranks + row * ld_out;
    const bool has_nan = row_nan[row] != 0;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < N;
         p += (long)gridDim.x * blockDim.x) {
        const int32_t cell = cs[p];
        if (has_nan) {
            out[cell] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        const uint32_t k = ks[p];
        const bool tie_left = p > 0 && ks[p - 1] == k;
        const bool tie_right = p + 1 < N && ks[p + 1] == k;
        if (!tie_left && !tie_right) {
            out[cell] = (double)(p + 1);
            continue;
        }
        long a = p, b = p + 1;                       // the bin [a, b): a short walk, then bisection
        {
            int steps = 0;
            while (a > 0 && ks[a - 1] == k && steps < kShortBin) { --a; ++steps; }
            if (a > 0 && ks[a - 1] == k) {
                long lo = 0, hi = a;                 // first index with ks[idx] == k
                while (lo < hi) {
                    const long mid = (lo + hi) >> 1;
                    if (ks[mid] < k) lo = mid + 1; else hi = mid;
                }
                a = lo;
            }
            steps = 0;
            while (b < N && ks[b] == k && steps < kShortBin) { ++b; ++steps; }
            if (b < N && ks[b] == k) {
                long lo = b, hi = N;                 // first index with ks[idx] > k
                while (lo < hi) {
                    const long mid = (lo + hi) >> 1;
                    if (ks[mid] > k) hi = mid; else lo = mid + 1;
                }
                b = lo;
            }
        }
        if (b - a > kShortBin) {
            if (p == a) {
                const unsigned slot = atomicAdd(work_count, 1u);
                work[slot] = LongBin{(int)row, (int)a, (int)b, 0};      // row: batch-relative
            }
            continue;
        }
        const double v = src[cell];
        int less = 0, equal = 0;
        for (long q = a; q < b; ++q) {
            const double u = src[cs[q]];
            less += u < v;
            equal += u == v;
        }
        // "average": mean of ranks a+less+1 .. a+less+equal (exact); "min": the smallest of them
        out[cell] = method == FP_RANK_MIN ? (double)(a + less + 1) : (double)(a + less) + 0.5 * (double)(equal + 1);
    }
  }
}

// One WARP per listed bin.  `scores`, `sorted_cell`, `out_ranks` are the batch bases and
// LongBin::row is batch-relative.  An all-equal bin is one tie run; a mixed bin of up to
// kWarpBin cells is counted by the warp; larger mixed bins move on to the CTA-wide kernel.
__global__ void __launch_bounds__(256)
long_bins_warp_kernel(const double *__restrict__ scores, const int32_t *__restrict__ sorted_cell, long ld,
                      double *__restrict__ out_ranks, long ld_out, int method, const LongBin *__restrict__ work,
                      const unsigned *__restrict__ work_count, LongBin *__restrict__ big,
                      unsigned *__restrict__ big_count) {
    const unsigned n_work = *work_count;
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n_work; e += warps) {
        const LongBin w = work[e];
        const double *src = scores + (long)w.row * ld;
        const int32_t *cs = sorted_cell + (long)w.row * ld;
        double *out = out_ranks + (long)w.row * ld_out;
        const long a = w.a, L = w.b - w.a;
        double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
        for (long i = lane; i < L; i += 32) {
            const double v = src[cs[a + i]];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
        for (int o = 16; o; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (mn == mx) {
            const double r = method == FP_RANK_MIN ? (double)(a + 1) : (double)a + 0.5 * (double)(L + 1);
            for (long i = lane; i < L; i += 32) out[cs[a + i]] = r;
            continue;
        }
        if (L > kWarpBin) {
            if (lane == 0) big[atomicAdd(big_count, 1u)] = w;
            continue;
        }
        for (long i = lane; i < L; i += 32) {
            const int32_t cell = cs[a + i];
            const double v = src[cell];
            int less = 0, equal = 0;
            for (long q = 0; q < L; ++q) {            // the whole warp reads the same address
                const double u = src[cs[a + q]];
                less += u < v;
                equal += u == v;
            }
            out[cell] = method == FP_RANK_MIN ? (double)(a + less + 1)
                                              : (double)(a + less) + 0.5 * (double)(equal + 1);
        }
    }
}

// One CTA per large mixed bin (more than kWarpBin cells whose short keys coincide although their
// doubles differ: a row squeezed by an outlier, small non-zero values next to an expression
// column's zeros).  Counting every cell against the whole bin is O(L^2) -- 1e10 compares for a
// bin of 1e5 cells.  Instead the bin is re-quantised on ITS OWN range [min, max] into kSubBins
// sub-bins (histogram and prefix in shared memory, members scattered into a second index
// array), and a cell counts only against the members of its sub-bin: O(L * L / kSubBins) for
// spread-out values, and all-equal sub-bins (the zeros) are recognised without comparing.  Exact:
// the sub-key is monotone in the value, ties are still decided on the doubles.
constexpr int kSubBins = 4096;

// A sub-bin that is still large and mixed (zeros next to a few tiny values) becomes a work item of
// the NEXT round, which re-quantises it on its own, 4 096 times narrower, range: the host launches
// a fixed number of rounds that ping-pong between the two index arrays and the two work lists; the
// last round counts whatever is left.
__global__ void __launch_bounds__(256)
long_bins_cta_kernel(const double *__restrict__ scores, const int32_t *__restrict__ sorted_cell,
                     int32_t *__restrict__ scratch_cell, long ld, double *__restrict__ out_ranks, long ld_out,
                     int method, const LongBin *__restrict__ big, const unsigned *__restrict__ big_count,
                     LongBin *__restrict__ next, unsigned *__restrict__ next_count) {
    __shared__ uint32_t sub[kSubBins + 1];
    __shared__ uint32_t mixed[kSubBins / 32];            // bit k: sub-bin k holds different values
    __shared__ double s_mn[8], s_mx[8];
    __shared__ uint32_t s_part[8];
    const unsigned n_big = *big_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned e = blockIdx.x; e < n_big; e += gridDim.x) {
        const LongBin w = big[e];
        const double *src = scores + (long)w.row * ld;
        const int32_t *cs = sorted_cell + (long)w.row * ld;
        int32_t *cs2 = scratch_cell + (long)w.row * ld;
        double *out = out_ranks + (long)w.row * ld_out;
        const long a = w.a, L = w.b - w.a;
        // range of the bin (finite or infinite: bit-pattern sub-keys then)
        double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
        for (long i = threadIdx.x; i < L; i += blockDim.x) {
            const double v = src[cs[a + i]];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
        for (int o = 16; o; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
        for (int k = threadIdx.x; k <= kSubBins; k += blockDim.x) sub[k] = 0;
        for (int k = threadIdx.x; k < kSubBins / 32; k += blockDim.x) mixed[k] = 0;
        __syncthreads();
        for (int k = 0; k < 8; ++k) { mn = fmin(mn, s_mn[k]); mx = fmax(mx, s_mx[k]); }
        const double span = mx - mn;
        const double scale = (span > 0.0 && !isinf(span)) ? (double)(kSubBins - 1) / span : 0.0;
        const bool linear = scale > 0.0 && !isinf(scale);
        // with bit-pattern keys the leading bits that min and max share carry no information
        const uint64_t bmn = ordered_bits(mn), bmx = ordered_bits(mx);
        const int shared_bits = (bmn == bmx) ? 52 : __clzll((long long)(bmn ^ bmx));
        const int shift = 64 - 12 - (shared_bits < 52 ? shared_bits : 52);
        auto subkey = [&](double v) -> uint32_t {
            if (linear) return (uint32_t)__double2ull_rz(fmin((v - mn) * scale, (double)(kSubBins - 1)));
            const uint64_t rel = ordered_bits(v) - bmn;             // monotone, 0 at the minimum
            const uint64_t q = shift > 0 ? (rel >> shift) : rel;
            return (uint32_t)(q < (uint64_t)(kSubBins - 1) ? q : (uint64_t)(kSubBins - 1));
        };
        for (long i = threadIdx.x; i < L; i += blockDim.x) atomicAdd(&sub[subkey(src[cs[a + i]])], 1u);
        __syncthreads();
        {   // exclusive prefix over the 4 096 sub-bin counts: warp w owns sub-bins 512 w .. 512 w + 511,
            // 32 consecutive ones per step (lane = counter: no bank conflicts)
            uint32_t *chunk = sub + warp * 512;
            uint32_t total = 0;
            for (int step = 0; step < 16; ++step) total += chunk[step * 32 + lane];
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) s_part[warp] = total;
            __syncthreads();
            uint32_t run = 0;
            for (int k = 0; k < warp; ++k) run += s_part[k];
            for (int step = 0; step < 16; ++step) {
                const uint32_t c = chunk[step * 32 + lane];
                uint32_t incl = c;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                chunk[step * 32 + lane] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (threadIdx.x == blockDim.x - 1) sub[kSubBins] = run;
        }
        __syncthreads();
        // members of every sub-bin, contiguous in the scratch index array; afterwards sub[k] = END of sub-bin k
        for (long i = threadIdx.x; i < L; i += blockDim.x) {
            const int32_t cell = cs[a + i];
            cs2[a + atomicAdd(&sub[subkey(src[cell])], 1u)] = cell;
        }
        __syncthreads();
        // a sub-bin whose members all equal its first one (the zeros of an expression column) is one tie run
        for (long i = threadIdx.x; i < L; i += blockDim.x) {
            const double v = src[cs[a + i]];
            const uint32_t k = subkey(v);
            const double first = src[cs2[a + (k ? sub[k - 1] : 0u)]];
            if (v != first) atomicOr(&mixed[k >> 5], 1u << (k & 31));
        }
        __syncthreads();
        for (long i = threadIdx.x; i < L; i += blockDim.x) {
            const int32_t cell = cs[a + i];
            const double v = src[cell];
            const uint32_t k = subkey(v);
            const uint32_t lo = k ? sub[k - 1] : 0u, hi = sub[k];
            long less = 0, equal = hi - lo;
            if ((mixed[k >> 5] >> (k & 31)) & 1u) {
                if (next && hi - lo > (uint32_t)kWarpBin && hi - lo < (uint32_t)L) {
                    // still large and mixed: one member files it for the next round (which reads
                    // the members from the array this round wrote)
                    if (cs2[a + lo] == cell) next[atomicAdd(next_count, 1u)] = LongBin{w.row, (int)(a + lo), (int)(a + hi), 0};
                    continue;
                }
                equal = 0;
                for (uint32_t q = lo; q < hi; ++q) {
                    const double u = src[cs2[a + q]];
                    less += u < v;
                    equal += u == v;
                }
            }
            const long below = a + lo + less;
            out[cell] = method == FP_RANK_MIN ? (double)(below + 1) : (double)below + 0.5 * (double)(equal + 1);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Rows of up to kFusedMaxCells cells: the whole ranking of a row inside ONE CTA, no sort at all.
// A 16-bit monotone key has 65 536 bins; their 16-bit counters (128 KB) and the row's cells in bin
// order (2 bytes each) fit the shared memory of an SM.  Histogram -> exclusive prefix (= where each
// bin starts) -> scatter of the cell numbers into bin order (= a counting sort; the order inside a
// bin does not matter) -> every cell reads its bin [start, end): alone: rank = start + 1; a few
// cells: count the bin's doubles below / equal to its own, exactly as assign_ranks_kernel does;
// more than kShortBin: the bin goes to the same device work list, and the row's bin order is copied
// out for the long-bin kernels.  Ranks are written in cell order (coalesced).  The row is read four
// times, three of them from L2.  Replaces quantise_rows_kernel + three CUB radix passes +
// assign_ranks_kernel (3.2 ms at 3 500 x 20 000) for these rows.
// ---------------------------------------------------------------------------------------------
constexpr int kFusedThreads = 1024;
constexpr int kFusedBins = 65536;
constexpr long kFusedMaxCells = 48000;             // 128 KB of counters + 2 bytes per cell <= 227 KB
constexpr long kFusedSubMaxCells = 24500;          // ... + 2 more bytes per cell for the sub-keys
static_assert(kFusedBins * 2 + kFusedMaxCells * 2 + 1024 <= 232448, "shared memory of rank_rows_fused_kernel");
static_assert(kFusedBins * 2 + kFusedSubMaxCells * 4 + 1024 <= 232448, "shared memory of rank_rows_fused_kernel<true>");

// 32-bit monotone key: the top 16 bits choose the bin, the low 16 bits ("sub-key") order most cells
// inside a bin without looking at their doubles
__device__ __forceinline__ uint32_t fused_key32(double v, int mode, double lo, double scale) {
    if (mode == 0) return (uint32_t)__double2ull_rz(fmin((v - lo) * scale, 4294967295.0));
    return (uint32_t)(ordered_bits(v) >> 32);
}
__device__ __forceinline__ uint32_t half_of(const uint32_t *words, uint32_t k) {
    return (words[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
}

// SUB: keep the sub-keys of the cells (in bin order) in shared memory.  A cell then decides
// "below / above" against a bin mate from the sub-keys alone -- the key is monotone, so a smaller
// sub-key means a smaller double -- and loads the mate's double only when the sub-keys coincide
// (true ties, or values closer than 2^-32 of the row's range).  Without it every cell of a shared
// bin (a quarter of the cells at 20 000 cells) issues dependent loads of its mates' doubles, and
// those chains, not bandwidth, set the time of the kernel (1.35 ms at 3 500 x 20 000).
template <bool SUB>
__global__ void __launch_bounds__(kFusedThreads, 1)
rank_rows_fused_kernel(const double *__restrict__ scores, long rows, long N, long ld, double *__restrict__ out_ranks,
                       long ld_out, int method, int32_t *__restrict__ cell_sorted, LongBin *__restrict__ work,
                       unsigned *__restrict__ work_count) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(fused_smem);                       // kFusedBins / 2 words
    uint16_t *order = reinterpret_cast<uint16_t *>(fused_smem + kFusedBins * 2);     // N cells in bin order
    uint16_t *subk = order + ((N + 7) & ~7L);                                        // their sub-keys (SUB)
    __shared__ double s_min[32], s_max[32];
    __shared__ int s_flag[32];
    __shared__ uint32_t s_part[32];
    __shared__ double s_lo, s_scale;
    __shared__ int s_mode, s_long;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long row = blockIdx.x; row < rows; row += gridDim.x) {
        const double *src = scores + row * ld;
        double *out = out_ranks + row * ld_out;
        // ---- range and special values (the only read of the row that misses L2)
        double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
        int flag = 0;                                        // bit 0: NaN seen, bit 1: infinity seen
        {
            auto see = [&](double v) {
                if (v != v) flag |= 1;
                else {
                    if (isinf(v)) flag |= 2;
                    mn = fmin(mn, v);
                    mx = fmax(mx, v);
                }
            };
            long c = tid;
            for (; c + 7 * kFusedThreads < N; c += 8 * kFusedThreads) {       // eight loads in flight (HBM)
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = src[c + u * kFusedThreads];
#pragma unroll
                for (int u = 0; u < 8; ++u) see(v[u]);
            }
            for (; c < N; c += kFusedThreads) see(src[c]);
        }
        for (int o = 16; o; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            flag |= __shfl_xor_sync(0xffffffffu, flag, o);
        }
        if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; s_flag[warp] = flag; }
        for (int w = tid; w < kFusedBins / 2; w += kFusedThreads) cnt[w] = 0;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kFusedThreads / 32; ++w) {
                mn = fmin(mn, s_min[w]);
                mx = fmax(mx, s_max[w]);
                flag |= s_flag[w];
            }
            const double span = mx - mn;                     // may overflow to +inf: bit-pattern keys
            const double scale = (span > 0.0 && !isinf(span)) ? 4294967295.0 / span : 0.0;
            s_lo = mn;
            s_scale = scale;                                 // a denormal span overflows it: bit-pattern keys
            s_mode = (flag & 1) ? 2 : ((flag & 2) || isinf(span) || isinf(scale)) ? 1 : 0;
            s_long = 0;
        }
        __syncthreads();
        const double lo = s_lo, scale = s_scale;
        const int mode = s_mode;
        if (mode == 2) {                                     // a NaN anywhere: NaN everywhere (rankdata propagates)
            for (long c = tid; c < N; c += kFusedThreads) out[c] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();
            continue;
        }
        // ---- histogram of the 16-bit bin numbers (two counters per word)
        {
            long c = tid;
            for (; c + 3 * kFusedThreads < N; c += 4 * kFusedThreads) {       // four loads in flight
                const double v0 = src[c], v1 = src[c + kFusedThreads], v2 = src[c + 2 * kFusedThreads],
                             v3 = src[c + 3 * kFusedThreads];
                const uint32_t q0 = fused_key32(v0, mode, lo, scale) >> 16, q1 = fused_key32(v1, mode, lo, scale) >> 16,
                               q2 = fused_key32(v2, mode, lo, scale) >> 16, q3 = fused_key32(v3, mode, lo, scale) >> 16;
                atomicAdd(&cnt[q0 >> 1], (q0 & 1) ? 0x10000u : 1u);
                atomicAdd(&cnt[q1 >> 1], (q1 & 1) ? 0x10000u : 1u);
                atomicAdd(&cnt[q2 >> 1], (q2 & 1) ? 0x10000u : 1u);
                atomicAdd(&cnt[q3 >> 1], (q3 & 1) ? 0x10000u : 1u);
            }
            for (; c < N; c += kFusedThreads) {
                const uint32_t q = fused_key32(src[c], mode, lo, scale) >> 16;
                atomicAdd(&cnt[q >> 1], (q & 1) ? 0x10000u : 1u);
            }
        }
        __syncthreads();
        // ---- exclusive prefix over the 65 536 counters.  Warp w owns words 1024 w .. 1024 w + 1023
        // and walks them 32 consecutive words at a time (lane = word: no bank conflicts; a thread
        // owning 32 consecutive words of its own would hit one bank with all 32 lanes -- measured
        // with ncu: 70 % of the shared-memory wavefronts of the first version were such conflicts).
        {
            uint32_t *chunk = cnt + warp * 1024;
            uint32_t total = 0;
#pragma unroll 4
            for (int step = 0; step < 32; ++step) {
                const uint32_t word = chunk[step * 32 + lane];
                total += (word & 0xffffu) + (word >> 16);
            }
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) s_part[warp] = total;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = s_part[lane], acc = v;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, acc, o);
                    if (lane >= o) acc += up;
                }
                s_part[lane] = acc - v;                      // exclusive over the warps
            }
            __syncthreads();
            uint32_t run = s_part[warp];                     // first position of this warp's first bin
            for (int step = 0; step < 32; ++step) {
                const uint32_t word = chunk[step * 32 + lane];
                const uint32_t c0 = word & 0xffffu, c1 = word >> 16, both = c0 + c1;
                uint32_t incl = both;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const uint32_t start = run + incl - both;    // first position of this word's even bin
                chunk[step * 32 + lane] = start | ((start + c0) << 16);      // positions < 65 536
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        // ---- counting sort: the cell numbers in bin order; afterwards a counter holds its bin's END
        {
            auto place = [&](long c, double v) {
                const uint32_t q32 = fused_key32(v, mode, lo, scale);
                const uint32_t q = q32 >> 16;
                const uint32_t old = atomicAdd(&cnt[q >> 1], (q & 1) ? 0x10000u : 1u);
                const uint32_t pos = (old >> ((q & 1) * 16)) & 0xffffu;
                order[pos] = (uint16_t)c;
                if (SUB) subk[pos] = (uint16_t)q32;
            };
            long c = tid;
            for (; c + 3 * kFusedThreads < N; c += 4 * kFusedThreads) {
                const double v0 = src[c], v1 = src[c + kFusedThreads], v2 = src[c + 2 * kFusedThreads],
                             v3 = src[c + 3 * kFusedThreads];
                place(c, v0);
                place(c + kFusedThreads, v1);
                place(c + 2 * kFusedThreads, v2);
                place(c + 3 * kFusedThreads, v3);
            }
            for (; c < N; c += kFusedThreads) place(c, src[c]);
        }
        __syncthreads();
        // ---- ranks, written in cell order (four cells per thread in flight)
        {
            auto rank_cell = [&](long c, double v) {
                const uint32_t q32 = fused_key32(v, mode, lo, scale);
                const uint32_t q = q32 >> 16;
                const uint32_t b = half_of(cnt, q);
                const uint32_t a = q ? half_of(cnt, q - 1) : 0u;  // the previous bin's end
                const uint32_t L = b - a;
                if (L == 1) {
                    out[c] = (double)(a + 1);
                } else if (L <= (uint32_t)kShortBin) {
                    int less = 0, equal = 0;
                    const uint32_t mine = q32 & 0xffffu;
                    for (uint32_t i = a; i < b; ++i) {
                        if (SUB) {
                            const uint32_t theirs = subk[i];
                            if (theirs != mine) {                 // monotone key: the sub-keys decide
                                less += theirs < mine;
                                continue;
                            }
                            if (order[i] == (uint16_t)c) {        // myself
                                ++equal;
                                continue;
                            }
                        }
                        const double u = src[order[i]];
                        less += u < v;
                        equal += u == v;
                    }
                    out[c] = method == FP_RANK_MIN ? (double)(a + less + 1)
                                                   : (double)(a + less) + 0.5 * (double)(equal + 1);
                } else if (order[a] == (uint16_t)c) {            // one cell of a long bin files it
                    const unsigned slot = atomicAdd(work_count, 1u);
                    work[slot] = LongBin{(int)row, (int)a, (int)b, 0};
                    s_long = 1;
                }
            };
            long c = tid;
            for (; c + 3 * kFusedThreads < N; c += 4 * kFusedThreads) {
                const double v0 = src[c], v1 = src[c + kFusedThreads], v2 = src[c + 2 * kFusedThreads],
                             v3 = src[c + 3 * kFusedThreads];
                rank_cell(c, v0);
                rank_cell(c + kFusedThreads, v1);
                rank_cell(c + 2 * kFusedThreads, v2);
                rank_cell(c + 3 * kFusedThreads, v3);
            }
            for (; c < N; c += kFusedThreads) rank_cell(c, src[c]);
        }
        __syncthreads();
        if (s_long) {                                        // the long-bin kernels read the bin order from global memory
            int32_t *cs = cell_sorted + row * ld;
            for (long p = tid; p < N; p += kFusedThreads) cs[p] = (int32_t)order[p];
        }
        __syncthreads();
    }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

long batch_rows(long K, long ld) {
    long b = kMaxBatchElems / ld;
    if (b < 1) b = 1;
    if (b > K) b = K;
    return b;
}

// key width: ~8 empty bins per cell, in whole radix digits
int key_bits(long N) {
    int lg = 0;
    while ((1L << lg) < N) ++lg;
    int bits = (lg + 3 + kRadixBits - 1) / kRadixBits * kRadixBits;
    if (bits < 2 * kRadixBits) bits = 2 * kRadixBits;
    if (bits > 30) bits = 30;
    return bits;
}

size_t cub_temp_bytes(long rows, long N, long ld) {
    size_t bytes = 0;
    CountIt cnt(0);
    BeginIt beg(cnt, RowBegin{ld});
    EndIt end(cnt, RowEnd{ld, N});
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr,
                                             (uint32_t *)nullptr, (const int32_t *)nullptr,
                                             (int32_t *)nullptr, (int)(rows * ld), (int)rows, beg,
                                             end, 0, key_bits(N), (cudaStream_t)0);
    return bytes;
}

struct Layout {
    size_t idx_b, nan_b, work_b, temp_b, total;
};

Layout layout(long rows, long N, long ld) {
    Layout l;
    l.idx_b = align_up((size_t)rows * ld * 4);
    l.nan_b = align_up((size_t)rows * 4);
    // at most one entry per kShortBin + 1 cells, plus the counter
    l.work_b = align_up(((size_t)rows * ld / (kShortBin + 1) + 1) * sizeof(LongBin) + 256) +
               2 * align_up(((size_t)rows * ld / (kWarpBin + 1) + 1) * sizeof(LongBin));
    l.temp_b = align_up(cub_temp_bytes(rows, N, ld));
    l.total = 4 * l.idx_b + l.nan_b + l.work_b + l.temp_b;
    return l;
}

// The rounds of long_bins_cta_kernel: list A -> list B -> list A -> (last: no further list).  The index
// arrays swap roles with the lists: a round reads the members of its bins from the array the
// previous round wrote.
cudaError_t long_bin_rounds(const double *src, int32_t *idx_a, int32_t *idx_b, long ld, double *dst, int method,
                            LongBin *list_a, unsigned *count_a, LongBin *list_b, unsigned *count_b, cudaStream_t st) {
    constexpr int kRounds = 4;
    for (int r = 0; r < kRounds; ++r) {
        const bool last = r == kRounds - 1;
        if (!last) {
            cudaError_t err = cudaMemsetAsync(count_b, 0, sizeof(unsigned), st);
            if (err != cudaSuccess) return err;
        }
        long_bins_cta_kernel<<<fp::sm_count() * 4, 256, 0, st>>>(src, idx_a, idx_b, ld, dst, ld, method, list_a, count_a,
                                                                 last ? nullptr : list_b, last ? nullptr : count_b);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        fp::count_launch();
        int32_t *ti = idx_a; idx_a = idx_b; idx_b = ti;
        LongBin *tl = list_a; list_a = list_b; list_b = tl;
        unsigned *tc = count_a; count_a = count_b; count_b = tc;
    }
    return cudaSuccess;
}

}  // namespace

extern "C" size_t fp_rank_workspace_bytes(int64_t K, int64_t N, int64_t ld) {
    if (K <= 0 || N <= 0 || ld < N) return 0;
    // parts: keys in / out, cell index in / out, NaN flags, long-bin work list, CUB temp
    return layout(batch_rows(K, ld), N, ld).total + 1024;
}

extern "C" int fp_rank_columns_ex(const double *scores, int64_t K, int64_t N, int64_t ld, double *out_ranks,
                                  void *workspace, size_t workspace_bytes, int method, void *stream);

extern "C" int fp_rank_columns(const double *scores, int64_t K, int64_t N, int64_t ld,
                               double *out_ranks, void *workspace, size_t workspace_bytes,
                               void *stream) {
    return fp_rank_columns_ex(scores, K, N, ld, out_ranks, workspace, workspace_bytes, FP_RANK_AVERAGE, stream);
}

extern "C" int fp_rank_columns_ex(const double *scores, int64_t K, int64_t N, int64_t ld, double *out_ranks,
                                  void *workspace, size_t workspace_bytes, int method, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(method == FP_RANK_AVERAGE || method == FP_RANK_MIN, "fp_rank_columns: unknown method %d", method);
    FP_CHECK_ARG(scores && out_ranks, "fp_rank_columns: null pointer");
    FP_CHECK_ARG(K >= 0 && N > 0 && ld >= N, "fp_rank_columns: bad shape K=%ld N=%ld ld=%ld",
                 (long)K, (long)N, (long)ld);
    FP_CHECK_ARG(N < (1L << 31), "fp_rank_columns: N too large");
    if (K == 0) return FP_OK;
    const long rows = batch_rows(K, ld);
    const Layout l = layout(rows, N, ld);
    if (!workspace || workspace_bytes < l.total) {
        fp::set_error("fp_rank_columns: workspace %zu bytes < required %zu", workspace_bytes, l.total);
        return FP_EWORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    uint32_t *keys_in = reinterpret_cast<uint32_t *>(ws);
    uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(ws + l.idx_b);
    int32_t *cell_in = reinterpret_cast<int32_t *>(ws + 2 * l.idx_b);
    int32_t *cell_sorted = reinterpret_cast<int32_t *>(ws + 3 * l.idx_b);
    int32_t *row_nan = reinterpret_cast<int32_t *>(ws + 4 * l.idx_b);
    unsigned *work_count = reinterpret_cast<unsigned *>(ws + 4 * l.idx_b + l.nan_b);
    LongBin *work = reinterpret_cast<LongBin *>(ws + 4 * l.idx_b + l.nan_b + 256);
    LongBin *big = reinterpret_cast<LongBin *>(
        ws + 4 * l.idx_b + l.nan_b + align_up(((size_t)rows * ld / (kShortBin + 1) + 1) * sizeof(LongBin) + 256));
    unsigned *big_count = work_count + 1;
    LongBin *big2 = reinterpret_cast<LongBin *>(reinterpret_cast<char *>(big) +
                                                align_up(((size_t)rows * ld / (kWarpBin + 1) + 1) * sizeof(LongBin)));
    unsigned *big2_count = work_count + 2;
    void *cub_temp = ws + 4 * l.idx_b + l.nan_b + l.work_b;
    cudaStream_t st = fp::as_stream(stream);
    const int bits = key_bits(N);

    const char *force = getenv("FASTPROJECT_B200_RANK");            // "sort": the CUB path for every N (tests)
    const bool fused = N <= kFusedMaxCells && !(force && !strcmp(force, "sort"));
    const bool with_sub = N <= kFusedSubMaxCells;
    const size_t fused_smem = (size_t)kFusedBins * 2 + (size_t)((N + 7) & ~7L) * (with_sub ? 4 : 2) + 16;
    if (fused) {
        static bool attr_set = false;
        if (!attr_set) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(rank_rows_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kFusedBins * 2 + (int)kFusedSubMaxCells * 4 + 64));
            FP_CHECK_CUDA(cudaFuncSetAttribute(rank_rows_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kFusedBins * 2 + (int)kFusedMaxCells * 2 + 64));
            attr_set = true;
        }
    }
    for (long k0 = 0; k0 < K; k0 += rows) {
        const long nb = (K - k0 < rows) ? (K - k0) : rows;
        const double *src = scores + k0 * ld;
        double *dst = out_ranks + k0 * ld;
        FP_CHECK_CUDA(cudaMemsetAsync(work_count, 0, 4 * sizeof(unsigned), st));
        if (fused) {
            const long ctas = nb < fp::sm_count() ? nb : fp::sm_count();
            if (with_sub)
                rank_rows_fused_kernel<true><<<(unsigned)ctas, kFusedThreads, fused_smem, st>>>(
                    src, nb, N, ld, dst, ld, method, cell_sorted, work, work_count);
            else
                rank_rows_fused_kernel<false><<<(unsigned)ctas, kFusedThreads, fused_smem, st>>>(
                    src, nb, N, ld, dst, ld, method, cell_sorted, work, work_count);
            FP_CHECK_LAUNCH("rank_rows_fused_kernel");
            long_bins_warp_kernel<<<fp::sm_count() * 8, 256, 0, st>>>(src, cell_sorted, ld, dst, ld, method, work,
                                                                      work_count, big, big_count);
            FP_CHECK_LAUNCH("long_bins_warp_kernel");
            FP_CHECK_CUDA(long_bin_rounds(src, cell_sorted, cell_in, ld, dst, method, big, big_count, big2, big2_count, st));
            continue;
        }
        const long cap = (long)fp::sm_count() * 4;
        quantise_rows_kernel<<<(unsigned)(nb < cap ? nb : cap), QUANT_THREADS, 0, st>>>(src, nb, N, ld, bits, keys_in,
                                                                                       cell_in, row_nan);
        FP_CHECK_LAUNCH("quantise_rows_kernel");
        CountIt cnt(0);
        BeginIt beg(cnt, RowBegin{ld});
        EndIt end(cnt, RowEnd{ld, N});
        size_t tb = l.temp_b;
        FP_CHECK_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(
            cub_temp, tb, keys_in, keys_sorted, cell_in, cell_sorted, (int)(nb * ld), (int)nb, beg, end, 0, bits,
            st));
        fp::count_launch(bits / kRadixBits);  // library launches, one per radix pass
        unsigned gx = (unsigned)((N + 255) / 256);
        if (gx > 64) gx = 64;
        const unsigned gy = (unsigned)(nb < 32768 ? nb : 32768);
        assign_ranks_kernel<<<dim3(gx, gy), 256, 0, st>>>(src, keys_sorted, cell_sorted, row_nan, nb, N, ld, dst, ld,
                                                          method, work, work_count);
        FP_CHECK_LAUNCH("assign_ranks_kernel");
        long_bins_warp_kernel<<<fp::sm_count() * 8, 256, 0, st>>>(src, cell_sorted, ld, dst, ld, method, work,
                                                                  work_count, big, big_count);
        FP_CHECK_LAUNCH("long_bins_warp_kernel");
        FP_CHECK_CUDA(long_bin_rounds(src, cell_sorted, cell_in, ld, dst, method, big, big_count, big2, big2_count, st));
    }
    return FP_OK;
}
