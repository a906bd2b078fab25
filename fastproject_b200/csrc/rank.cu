// Average ranks of every signature's scores over the cells.
//
// Replaces SignatureScores.ranks = scipy.stats.rankdata(scores, "average")
// (/root/reference/FastProject/SigScoreMethods.py:170-178) for all K signatures
// at once, and the rank-matrix assembly of Signatures.py:359-369.
//
// Step 1: segmented LSD radix sort of (score, cell) pairs, one segment per
//         signature row (CUB device primitive -- a library call; this step is
//         sort-bound and far below the contraction in cost).
// Step 2: our kernel turns sorted positions into 1-based average ranks:
//         a key without an equal neighbour gets position+1; members of a tie
//         run [a,b) all get (a + b + 1) / 2, the run being found by two binary
//         searches on the sorted row (no serial walk, so a constant row costs
//         the same as any other).  Comparisons are on double values, so -0.0
//         ties with +0.0 exactly as in rankdata.  A row holding a NaN yields
//         NaN everywhere (scipy nan_policy="propagate").
#include <cub/device/device_segmented_radix_sort.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "fp_common.cuh"

namespace {

constexpr long kMaxBatchElems = 1L << 27;  // elements sorted per CUB call

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

__global__ void fill_cell_index_kernel(int32_t *__restrict__ idx, long rows, long N, long ld) {
    const long total = rows * ld;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        const long c = i % ld;
        idx[i] = (int32_t)(c < N ? c : 0);
    }
}

__global__ void __launch_bounds__(256)
assign_average_ranks_kernel(const double *__restrict__ sorted_keys,
                            const int32_t *__restrict__ sorted_cell, long rows, long N, long ld,
                            double *__restrict__ out_ranks, long ld_out, int method) {
  for (long row = blockIdx.y; row < rows; row += gridDim.y) {
    const double *ks = sorted_keys + row * ld;
    const int32_t *cs = sorted_cell + row * ld;
    double *out = out_ranks + row * ld_out;
    const double first = ks[0], last = ks[N - 1];
    const bool has_nan = (first != first) || (last != last);
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < N;
         p += (long)gridDim.x * blockDim.x) {
        const double v = ks[p];
        double rank;
        if (has_nan) {
            rank = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const bool tie_left = p > 0 && ks[p - 1] == v;
            const bool tie_right = p + 1 < N && ks[p + 1] == v;
            if (!tie_left && !tie_right) {
                rank = (double)(p + 1);
            } else {
                long lo = 0, hi = p;                 // first index with ks[idx] == v
                while (lo < hi) {
                    const long mid = (lo + hi) >> 1;
                    if (ks[mid] < v) lo = mid + 1; else hi = mid;
                }
                const long a = lo;
                lo = p + 1; hi = N;                  // first index with ks[idx] > v
                while (lo < hi) {
                    const long mid = (lo + hi) >> 1;
                    if (ks[mid] > v) hi = mid; else lo = mid + 1;
                }
                const long b = lo;
                // "average": mean of ranks a+1 .. b (exact); "min": the smallest of them
                rank = method == FP_RANK_MIN ? (double)(a + 1) : 0.5 * (double)(a + b + 1);
            }
        }
        out[cs[p]] = rank;
    }
  }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

long batch_rows(long K, long ld) {
    long b = kMaxBatchElems / ld;
    if (b < 1) b = 1;
    if (b > K) b = K;
    return b;
}

size_t cub_temp_bytes(long rows, long N, long ld) {
    size_t bytes = 0;
    CountIt cnt(0);
    BeginIt beg(cnt, RowBegin{ld});
    EndIt end(cnt, RowEnd{ld, N});
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, (const double *)nullptr,
                                             (double *)nullptr, (const int32_t *)nullptr,
                                             (int32_t *)nullptr, (int)(rows * ld), (int)rows, beg,
                                             end, 0, 64, (cudaStream_t)0);
    return bytes;
}

}  // namespace

extern "C" size_t fp_rank_workspace_bytes(int64_t K, int64_t N, int64_t ld) {
    if (K <= 0 || N <= 0 || ld < N) return 0;
    const long rows = batch_rows(K, ld);
    // parts: sorted keys, cell index in, cell index out, CUB temp
    return align_up((size_t)rows * ld * 8) + 2 * align_up((size_t)rows * ld * 4) +
           align_up(cub_temp_bytes(rows, N, ld)) + 1024;
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
    const size_t keys_b = align_up((size_t)rows * ld * 8), idx_b = align_up((size_t)rows * ld * 4);
    const size_t temp_b = cub_temp_bytes(rows, N, ld);
    if (!workspace || workspace_bytes < keys_b + 2 * idx_b + align_up(temp_b)) {
        fp::set_error("fp_rank_columns: workspace %zu bytes < required %zu", workspace_bytes,
                      keys_b + 2 * idx_b + align_up(temp_b));
        return FP_EWORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    double *keys_sorted = reinterpret_cast<double *>(ws);
    int32_t *cell_in = reinterpret_cast<int32_t *>(ws + keys_b);
    int32_t *cell_sorted = reinterpret_cast<int32_t *>(ws + keys_b + idx_b);
    void *cub_temp = ws + keys_b + 2 * idx_b;
    cudaStream_t st = fp::as_stream(stream);

    fill_cell_index_kernel<<<fp::sm_count() * 8, 256, 0, st>>>(cell_in, rows, N, ld);
    FP_CHECK_LAUNCH("fill_cell_index_kernel");

    for (long k0 = 0; k0 < K; k0 += rows) {
        const long nb = (K - k0 < rows) ? (K - k0) : rows;
        CountIt cnt(0);
        BeginIt beg(cnt, RowBegin{ld});
        EndIt end(cnt, RowEnd{ld, N});
        size_t tb = temp_b;
        FP_CHECK_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(
            cub_temp, tb, scores + k0 * ld, keys_sorted, cell_in, cell_sorted, (int)(nb * ld),
            (int)nb, beg, end, 0, 64, st));
        fp::count_launch(8);  // library launches (one per radix pass), counted approximately
        unsigned gx = (unsigned)((N + 255) / 256);
        if (gx > 64) gx = 64;
        const unsigned gy = (unsigned)(nb < 32768 ? nb : 32768);
        assign_average_ranks_kernel<<<dim3(gx, gy), 256, 0, st>>>(
            keys_sorted, cell_sorted, nb, N, ld, out_ranks + k0 * ld, ld, method);
        FP_CHECK_LAUNCH("assign_average_ranks_kernel");
    }
    return FP_OK;
}
