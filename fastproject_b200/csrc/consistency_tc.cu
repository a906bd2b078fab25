// Signature x projection local dissimilarity on the Blackwell tensor cores.
//
// Replaces, for one projection, /root/reference/FastProject/Signatures.py:396-408
// and :412-413 (cdist -> exp(-d^2/sigma^2) -> zero diagonal -> row normalise ->
// np.dot(weights, rank matrix) -> |rank - prediction|) for real and random
// signatures together, WITHOUT materialising the N x N kernel matrix.
//
// Arithmetic (FP_PRECISION_TC): an error-free digit-sliced contraction on
// tcgen05.mma kind::i8 (s32 accumulators in tensor memory -- integer
// accumulation is exact, so no tensor-core rounding enters the result):
//
//   prediction_i = sum_j w_ij v_j / sum_j w_ij,  w_ij = exp(-(d_ij^2 - m_i)/sigma_i^2)
//
//   * m_i = min_{j != i} d_ij^2 (nearest neighbour): a per-row scale that cancels
//     in the quotient and puts the largest weight of every row at exactly 1;
//   * W_ij = round(w_ij * (2^32 - 1)) is a 32-bit fixed-point weight, generated
//     in FP64 (table + cubic for 2^t, |error| < 2^-33) ONCE per projection by
//     weight_planes_kernel and stored as four unsigned 8-bit digit planes
//     [digit][cell i][cell j] (the UMMA B operand) -- 4 bytes per weight, 1.6 GB
//     at 20 000 cells, generated in slabs of cells i when N x N x 4 bytes exceeds
//     the slab budget (100 000 and 1 000 000 cells);
//   * the values are half-integers (average ranks, one-hot / binary columns), so
//     r_j = (2 v_j - centre) * 2^shift is an integer (centre and shift from the value
//     range: N + 1 for ranks; the top digit is always well used); it is cut into two
//     (N <= 32 640) or three balanced signed base-256 digits (planes of s8, the UMMA A
//     operand, loaded by TMA, 128-byte swizzle);
//   * every (weight digit a, value digit b) plane pair with a + b <= 3 is one MMA;
//     pairs of equal weight share an s32 accumulator ("class" a + b); sum_j W_ij r_j
//     is rebuilt exactly in int64 in the epilogue.  The classes below 2^-32 of the top
//     one (below the weight quantisation) are not computed; with three digits the
//     accumulators are drained every 16 384 cells j so that they cannot overflow;
//   * the normaliser sum_j W_ij comes from the same MMAs through a row of ones
//     appended to every block of 127 signatures, so numerator and denominator
//     use identical quantised weights (the quotient is an exact weighted mean
//     with weights perturbed by < 2^-33 relative to the row maximum).
//
// Tile (consistency_mma_kernel): a PAIR of CTAs (cluster of 2, tcgen05 cta_group::2) owns
// 2 x 128 signature rows (M = 256) x 128 cells i; j is streamed in stages of 128.  Each CTA
// loads by TMA (128-byte swizzle) the value planes of its own 128 rows and the weight planes
// of 64 of the 128 cells -- the MMA reads the B operand from both CTAs' shared memory.
// 4 classes x 128 columns = all 512 TMEM columns of each SM.  Warp roles: warp 0 TMA
// producer, warp 1 MMA issuer (leader CTA, one elected lane), warps 2-17 epilogue.  The
// launch puts the pairs of signature blocks of one cell tile next to each other, so the
// pairs in flight stream the same weight planes and all but the first read them from L2.
// (Until round 1's last sessions the weights were generated inside this kernel by 16
// generator warps and never left shared memory; every pair of signature blocks then
// regenerated the whole N x N matrix -- 14 times at 3 500 signatures -- and the FP64 generator,
// not the tensor pipe, bounded the kernel: 14.0 ms per projection against 7.5 ms now.)
// The FP64 verifier lives in consistency_fp64.cu.
//
// Environment: FASTPROJECT_B200_SLAB_BYTES (weight-plane budget, tests use it to run many
// slabs at small N).  No run-time switch changes the arithmetic: the "loads only" timing
// experiment of round 1 is a compile-time option (-DFP_TC_TIMING_NO_MMA; its round-2 twins:
// -DFP_TC_TIMING_NO_LOADS issues the MMAs without operand traffic, -DFP_TC_TIMING_SAME_TILE makes
// every load an L2 hit, -DFP_TC_TIMING_TWO_STAGES shortens the pipeline, -DFP_TC_L2_PROMOTION=...
// changes the tensor maps' L2 promotion), never part of a release build
// (tests/cuda/mma_bound_probe.py builds them into tests/cuda/_bin/; results in DESIGN.md 6b).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "fp_common.cuh"
#include "row_tree.cuh"
#include "tc_common.cuh"

#ifndef FP_TC_L2_PROMOTION
#define FP_TC_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif

namespace {

using namespace fp::tc;

constexpr int TM = 128;            // signature rows per CTA (127 + the ones row)
constexpr int SIGS_PER_TILE = 127;
constexpr int TN = 128;            // cells i per CTA pair
constexpr int TNH = TN / 2;        // cells whose weight planes one CTA of the pair loads
constexpr int KS = 128;            // cells j per stage (bytes of K per digit plane)
constexpr int WD = 4;              // weight digits (32-bit fixed point)
constexpr int EPI_WARPS = 16;      // 4 per TMEM lane quadrant
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int A_PLANE_BYTES = TM * KS;    // 16 KB: one value digit plane, one stage
constexpr int B_PLANE_BYTES = TNH * KS;   //  8 KB: one weight digit plane of this CTA's 64 cells
constexpr int NCLASS = 4;
constexpr int TAB = 128;           // 2^(k/128) table ...
constexpr int TAB_COPIES = 16;     // ... 16 interleaved copies (lanes l and l+16 share one; 32 copies of a
                                   // 64-entry table were measured: no faster)
// s32 accumulators: a class sums at most 3 digit products of magnitude <= 255 * 128 per j
constexpr int CHUNK_STAGES_3 = 128;       // three value digits: drain every 16 384 cells j

// cell records prepared by cell_setup_kernel
struct __align__(32) CellI { double alpha, c, x2c, y2c; };   // row side
struct __align__(32) CellJ { double x, y, nb, pad; };         // column side

// ---------------------------------------------------------------------------
// per-cell set-up: nearest-neighbour distance (row scale) and the coefficients
// of t_ij = log2(e)/sigma_i^2 * (m_i - d_ij^2) = alpha_i + c_i*nb_j + X_i*x_j + Y_i*y_j
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dist_key(double v) {   // v >= 0: bit pattern is monotone
    return (unsigned long long)__double_as_longlong(v);
}
// order-preserving 64-bit key of any double, and back (ranges and bounding boxes by atomicMin / Max)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_key(unsigned long long key) {
    const unsigned long long u = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
    return __longlong_as_double((long long)u);
}

// nearest-neighbour squared distance of every cell; the j range is split over blockIdx.y and
// combined with atomicMin (non-negative doubles order like their bit patterns)
__global__ void __launch_bounds__(256)
nearest_neighbour_kernel(const double *__restrict__ coords, long N, unsigned long long *__restrict__ nn) {
    __shared__ double sx[256], sy[256];
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    const double *cx = coords, *cy = coords + N;
    const bool live = i < N;
    const double xi = live ? cx[i] : 0.0, yi = live ? cy[i] : 0.0;
    double m = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    const long jper = ((N + gridDim.y - 1) / gridDim.y + 255) / 256 * 256;
    const long jlo = (long)blockIdx.y * jper, jhi = min(N, jlo + jper);
    for (long j0 = jlo; j0 < jhi; j0 += 256) {
        const long jj = j0 + threadIdx.x;
        sx[threadIdx.x] = jj < N ? cx[jj] : 0.0;
        sy[threadIdx.x] = jj < N ? cy[jj] : 0.0;
        __syncthreads();
        const int lim = (int)((jhi - j0) < 256 ? (jhi - j0) : 256);
#pragma unroll 4
        for (int t = 0; t < lim; ++t) {
            const double dx = xi - sx[t], dy = yi - sy[t];
            const double d2 = fma(dx, dx, dy * dy);
            if (j0 + t != i && d2 < m) m = d2;
        }
        __syncthreads();
    }
    if (live) atomicMin(&nn[i], dist_key(m));
}

// ---------------------------------------------------------------------------
// The same minimum through a uniform grid (about one cell of the projection per grid cell):
// cells are counting-sorted by grid cell, and every cell searches the Chebyshev rings around
// its own grid cell until the best distance found is no larger than the distance to the border
// of the block searched so far.  The squared distances are the brute-force kernel's
// (fma(dx, dx, dy * dy) of the same differences), so the result is bit-identical to it; only
// the candidates differ.  O(N) for spread-out projections instead of O(N^2) (0.25 ms at
// 20 000 cells, 6 ms at 100 000, 0.6 s at 1 000 000 for the brute-force kernel).
// ---------------------------------------------------------------------------
struct NnGrid {
    double xmin, ymin, hx, hy, inv_hx, inv_hy;
};

// bounds[0..1]: ordered keys of min x, min y (initialised to all ones); [2..3]: max x, max y
// (zero); [4..7]: sums of x, x^2, y, y^2 as doubles (zero).  The grid covers mean +- 3 sigma of
// each axis (clipped to the extremes): an outlier cell a thousand bandwidths away would
// otherwise stretch the grid until the bulk of the cells shared a handful of grid cells.  Cells
// outside the box fall into the border grid cells, whose outer sides never bound the search,
// so the result does not depend on the box (nor on the order of the atomic sums).
__global__ void __launch_bounds__(256)
nn_bounds_kernel(const double *__restrict__ coords, long N, unsigned long long *__restrict__ bounds) {
    unsigned long long lo_x = ~0ull, lo_y = ~0ull, hi_x = 0ull, hi_y = 0ull;
    double sx = 0.0, sxx = 0.0, sy = 0.0, syy = 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long)gridDim.x * 256) {
        const double x = coords[i], y = coords[N + i];
        const unsigned long long kx = ordered_key(x), ky = ordered_key(y);
        lo_x = kx < lo_x ? kx : lo_x;
        hi_x = kx > hi_x ? kx : hi_x;
        lo_y = ky < lo_y ? ky : lo_y;
        hi_y = ky > hi_y ? ky : hi_y;
        sx += x; sxx += x * x; sy += y; syy += y * y;
    }
    for (int off = 16; off; off >>= 1) {
        unsigned long long o;
        o = __shfl_xor_sync(0xffffffffu, lo_x, off); lo_x = o < lo_x ? o : lo_x;
        o = __shfl_xor_sync(0xffffffffu, lo_y, off); lo_y = o < lo_y ? o : lo_y;
        o = __shfl_xor_sync(0xffffffffu, hi_x, off); hi_x = o > hi_x ? o : hi_x;
        o = __shfl_xor_sync(0xffffffffu, hi_y, off); hi_y = o > hi_y ? o : hi_y;
        sx += __shfl_xor_sync(0xffffffffu, sx, off);
        sxx += __shfl_xor_sync(0xffffffffu, sxx, off);
        sy += __shfl_xor_sync(0xffffffffu, sy, off);
        syy += __shfl_xor_sync(0xffffffffu, syy, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bounds[0], lo_x);
        atomicMin(&bounds[1], lo_y);
        atomicMax(&bounds[2], hi_x);
        atomicMax(&bounds[3], hi_y);
        double *sums = reinterpret_cast<double *>(bounds + 4);
        atomicAdd(&sums[0], sx);
        atomicAdd(&sums[1], sxx);
        atomicAdd(&sums[2], sy);
        atomicAdd(&sums[3], syy);
    }
}

__device__ __forceinline__ void nn_axis(double vmin, double vmax, double s, double ss, long N, int G, double *lo,
                                        double *h) {
    const double mean = s / (double)N, var = ss / (double)N - mean * mean;
    const double sd = var > 0.0 ? sqrt(var) : 0.0;
    double a = fmax(vmin, mean - 3.0 * sd), b = fmin(vmax, mean + 3.0 * sd);
    if (!(b > a)) {                                   // degenerate or non-finite: the extremes
        a = vmin;
        b = vmax;
    }
    const double w = b - a;
    *lo = a;
    *h = (w > 0.0 && w < 1.0e300) ? w / G : 1.0;
}

__device__ __forceinline__ NnGrid nn_grid_of(const unsigned long long *__restrict__ bounds, long N, int G) {
    const double *sums = reinterpret_cast<const double *>(bounds + 4);
    NnGrid g;
    nn_axis(from_ordered_key(bounds[0]), from_ordered_key(bounds[2]), sums[0], sums[1], N, G, &g.xmin, &g.hx);
    nn_axis(from_ordered_key(bounds[1]), from_ordered_key(bounds[3]), sums[2], sums[3], N, G, &g.ymin, &g.hy);
    g.inv_hx = 1.0 / g.hx;
    g.inv_hy = 1.0 / g.hy;
    return g;
}
__device__ __forceinline__ int nn_cell_coord(double v, double vmin, double inv_h, int G) {
    const double t = (v - vmin) * inv_h;
    int c = t >= (double)G ? G - 1 : (int)t;      // also catches the maximum (t == G)
    if (c < 0) c = 0;
    return c;
}

__global__ void __launch_bounds__(256)
nn_count_kernel(const double *__restrict__ coords, long N, const unsigned long long *__restrict__ bounds, int G,
                int *__restrict__ cell_of, int *__restrict__ counts) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const NnGrid g = nn_grid_of(bounds, N, G);
    const int c = nn_cell_coord(coords[N + i], g.ymin, g.inv_hy, G) * G + nn_cell_coord(coords[i], g.xmin, g.inv_hx, G);
    cell_of[i] = c;
    atomicAdd(&counts[c], 1);
}

__global__ void __launch_bounds__(256)
nn_scatter_kernel(const double *__restrict__ coords, long N, const int *__restrict__ cell_of,
                  const int *__restrict__ offsets, int *__restrict__ fill, double2 *__restrict__ sorted_xy,
                  int *__restrict__ sorted_id) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int c = cell_of[i];
    const int pos = offsets[c] + atomicAdd(&fill[c], 1);
    sorted_xy[pos] = make_double2(coords[i], coords[N + i]);
    sorted_id[pos] = (int)i;
}

__global__ void __launch_bounds__(128)
nn_search_kernel(const double2 *__restrict__ sorted_xy, const int *__restrict__ sorted_id, long N,
                 const int *__restrict__ offsets, const unsigned long long *__restrict__ bounds, int G,
                 unsigned long long *__restrict__ nn) {
    const long t = (long)blockIdx.x * 128 + threadIdx.x;
    if (t >= N) return;
    const NnGrid g = nn_grid_of(bounds, N, G);
    const double2 me = sorted_xy[t];
    const int pcx = nn_cell_coord(me.x, g.xmin, g.inv_hx, G), pcy = nn_cell_coord(me.y, g.ymin, g.inv_hy, G);
    double best = __longlong_as_double(0x7ff0000000000000LL);
    for (int r = 0;; ++r) {
        const int x0 = pcx - r, x1 = pcx + r, y0 = pcy - r, y1 = pcy + r;
        const int ya = y0 > 0 ? y0 : 0, yb = y1 < G - 1 ? y1 : G - 1;
        const int xa = x0 > 0 ? x0 : 0, xb = x1 < G - 1 ? x1 : G - 1;
        for (int uy = ya; uy <= yb; ++uy) {
            const bool full_row = uy == y0 || uy == y1;
            // a full row of the ring is one contiguous run of the sorted cells; inside rows
            // contribute their two end cells
            const int n_seg = full_row ? 1 : (r == 0 ? 1 : 2);
            for (int seg = 0; seg < n_seg; ++seg) {
                int ca, cb;
                if (full_row) {
                    ca = xa; cb = xb;
                } else {
                    const int ux = seg == 0 ? x0 : x1;
                    if (ux < 0 || ux > G - 1) continue;
                    ca = cb = ux;
                }
                const int qa = offsets[uy * G + ca], qb = offsets[uy * G + cb + 1];
                for (int q = qa; q < qb; ++q) {
                    if (q == t) continue;
                    const double2 o = sorted_xy[q];
                    const double dx = me.x - o.x, dy = me.y - o.y;
                    const double d2 = fma(dx, dx, dy * dy);
                    if (d2 < best) best = d2;
                }
            }
        }
        if (best == 0.0) break;
        // distance to the border of the block searched so far; sides on the edge of the grid do not count
        double bound = __longlong_as_double(0x7ff0000000000000LL);
        if (x0 > 0) bound = fmin(bound, me.x - (g.xmin + x0 * g.hx));
        if (x1 < G - 1) bound = fmin(bound, (g.xmin + (x1 + 1) * g.hx) - me.x);
        if (y0 > 0) bound = fmin(bound, me.y - (g.ymin + y0 * g.hy));
        if (y1 < G - 1) bound = fmin(bound, (g.ymin + (y1 + 1) * g.hy) - me.y);
        if (!(bound < __longlong_as_double(0x7ff0000000000000LL))) break;       // the whole grid was searched
        bound = fmax(bound, 0.0) * (1.0 - 1e-9);          // rounding of the cell borders
        if (best <= bound * bound) break;
    }
    nn[sorted_id[t]] = dist_key(best);
}

// ---------------------------------------------------------------------------
// Cell order of the contraction.  The sums over the cells j are exact integers, so the cells may
// be visited in any order; visiting them along a space-filling (Hilbert) curve makes every tile of 128
// consecutive cells spatially compact, and then whole (cell tile x 128-cell stage) blocks of the
// weight matrix have all-zero HIGH digit planes (pairs of cells further apart than 0.78 lose the top
// byte of the 32-bit weight at sigma = 0.33): weight_planes_kernel records which planes of a block
// are non-zero, and the MMA issuer skips the products of the all-zero ones -- exactly zero
// contributions, so the result does not change by a bit.
// ---------------------------------------------------------------------------
// position of grid point (x, y), 16 bits each, along the Hilbert curve of order 16 (the classic
// xy -> d walk: one quadrant per step, rotating the frame).  Hilbert rather than Morton order:
// consecutive cells are always neighbours, so tiles of 128 consecutive cells are more compact
// (measured on the host, tests/cuda/skip_estimate.py: 1 - 5 % fewer products left than with Z-order).
__device__ __forceinline__ uint32_t hilbert16(uint32_t x, uint32_t y) {
    uint32_t d = 0;
#pragma unroll
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = 65535u - x;
                y = 65535u - y;
            }
            const uint32_t t = x;
            x = y;
            y = t;
        }
    }
    return d;
}

__global__ void __launch_bounds__(256)
curve_key_kernel(const double *__restrict__ coords, long N, const unsigned long long *__restrict__ bounds,
                  uint32_t *__restrict__ keys, int32_t *__restrict__ cells) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const NnGrid g = nn_grid_of(bounds, N, 65536);          // the 3-sigma box of the cells, 65 536 steps a side
    const uint32_t qx = (uint32_t)nn_cell_coord(coords[i], g.xmin, g.inv_hx, 65536);
    const uint32_t qy = (uint32_t)nn_cell_coord(coords[N + i], g.ymin, g.inv_hy, 65536);
    keys[i] = hilbert16(qx, qy);
    cells[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
cell_setup_kernel(const double *__restrict__ coords, const double *__restrict__ sigma_sq,
                  double sigma_sq_scalar, long N, long Npad, const unsigned long long *__restrict__ nn,
                  const int32_t *__restrict__ order, CellI *__restrict__ ci, CellJ *__restrict__ cj,
                  double *__restrict__ table) {
    // 2^(k/128) * (2^32 - 1), 16 interleaved copies: every CTA of weight_planes_kernel copies it
    // into shared memory instead of evaluating 2 048 exp2 of its own
    if (blockIdx.x == 0)
        for (int idx = threadIdx.x; idx < TAB * TAB_COPIES; idx += 256)
            table[idx] = exp2((double)(idx / TAB_COPIES) * (1.0 / TAB)) * 4294967295.0;
    const long i = (long)blockIdx.x * 256 + threadIdx.x;       // position in the contraction's cell order
    if (i >= Npad) return;
    const double *cx = coords, *cy = coords + N;
    const bool live = i < N;
    const long cell = live ? (order ? (long)order[i] : i) : 0;
    const double xi = live ? cx[cell] : 0.0, yi = live ? cy[cell] : 0.0;
    const double m = live ? __longlong_as_double((long long)nn[cell]) : __longlong_as_double(0x7ff0000000000000LL);
    const double s2 = sigma_sq ? (live ? sigma_sq[cell] : 1.0) : sigma_sq_scalar;
    const double c = 1.4426950408889634074 / s2;             // log2(e) / sigma^2
    CellI a;
    a.c = c;
    a.x2c = 2.0 * c * xi;
    a.y2c = 2.0 * c * yi;
    // a row whose largest weight underflows to 0 in the reference (or that has no
    // other cell) has row sum 0 -> prediction 0 (Signatures.py:400-402)
    const bool dead = !live || !(m < __longlong_as_double(0x7ff0000000000000LL)) || exp(-m / s2) == 0.0;
    // - 2^-30: keeps t_ij <= 0 against the rounding of the expanded form (a common row
    // factor 2^(-2^-30), cancels in the quotient).  A dead row gets a large negative exponent
    // (all its weights round to 0 -> row sum 0 -> prediction 0).
    a.alpha = dead ? -1.0e5 : c * (m - fma(xi, xi, yi * yi)) - 9.313225746154785e-10;
    ci[i] = a;
    CellJ b;
    b.x = xi;
    b.y = yi;
    // padded columns need no special weight: their value digits (ones row included) are zero
    b.nb = -fma(xi, xi, yi * yi);
    b.pad = 0.0;
    cj[i] = b;
}

// ---------------------------------------------------------------------------
// value range -> centre and power-of-two scale of the digit representation.
// r = (2 v - centre2) * 2^shift is an integer whose top digit is always well
// used: for average ranks centre2 = N + 1; for one-hot / binary columns
// r = +-2^14, so the full 32-bit weight meets the only non-zero digit.
// range[0] = ordered key of min(v), range[1] = ordered key of max(v).
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
value_range_kernel(const double *__restrict__ values, long K, long N, long ld,
                   unsigned long long *__restrict__ range) {
    unsigned long long lo = ~0ull, hi = 0ull;
    bool nan = false;
    for (long row = blockIdx.y; row < K; row += gridDim.y)
        for (long j = (long)blockIdx.x * 256 + threadIdx.x; j < N; j += (long)gridDim.x * 256) {
            const double v = values[row * ld + j];
            nan |= (v != v);
            const unsigned long long k = ordered_key(v);
            lo = k < lo ? k : lo;
            hi = k > hi ? k : hi;
        }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, off);
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&range[0], lo);
        atomicMax(&range[1], hi);
    }
    if (nan) atomicMax(&range[2], 1ull);
}

struct Digits {
    double centre2;   // integer valued
    double scale;     // 2^shift
    bool valid;
};
template <int SR>
__device__ __forceinline__ Digits decode_range(const unsigned long long *__restrict__ range) {
    constexpr double LIMIT = SR == 2 ? 32639.0 : 8355711.0;   // 127 * (256^SR - 1) / 255
    const double mn2 = 2.0 * from_ordered_key(range[0]), mx2 = 2.0 * from_ordered_key(range[1]);
    Digits d;
    d.centre2 = floor(0.5 * (mn2 + mx2));
    // + 1: room for the parity adjustment of the centre below
    const double hr = fmax(mx2 - d.centre2, d.centre2 - mn2) + 1.0;
    d.valid = range[2] == 0ull && hr <= LIMIT;               // false for NaN / inf / too wide
    d.scale = 1.0;
    for (int it = 0; it < 24 && 2.0 * d.scale * hr <= LIMIT; ++it) d.scale *= 2.0;
    // The digit products of the lowest class are not computed, so the least significant digit
    // should average to zero over the cells.  For integer-spaced values (ranks without ties) all
    // r = 2 v - centre2 share one parity: with scale 128 even r make that digit 0; otherwise odd
    // r make it symmetric (e.g. +-64) where even r would give the lopsided set {0, -128}.
    const bool want_even = d.scale == 128.0;
    const bool is_even = fmod(mn2 - d.centre2, 2.0) == 0.0;
    if (want_even != is_even) d.centre2 += 1.0;
    return d;
}

// ---------------------------------------------------------------------------
// values -> balanced base-256 digit planes of r = (2 v - centre2) * scale, blocked as the
// A operand wants: plane b, tile row R = 128*blk + r holds signature 127*blk + r
// (r < 127) and the ones row (r = 127: most significant digit 1, others 0).
// planes[b][R][j], j padded with zeros to Npad.  Sets *flag if a value is not a
// half-integer or r does not fit SR digits.
// ---------------------------------------------------------------------------
template <int SR>
__global__ void __launch_bounds__(256)
pack_values_kernel(const double *__restrict__ values, long K, long N, long ld, long Kpad, long Npad,
                   const int32_t *__restrict__ order, int8_t *__restrict__ planes,
                   const unsigned long long *__restrict__ range, int *__restrict__ flag) {
    const Digits dg = decode_range<SR>(range);
    const long R = blockIdx.y;
    const long j4 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (j4 >= Npad) return;
    const long blk = R / TM, r = R % TM;
    const long sig = blk * SIGS_PER_TILE + r;
    uint32_t word[SR];
#pragma unroll
    for (int b = 0; b < SR; ++b) word[b] = 0;
    bool bad = !dg.valid;
    if (r == TM - 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (j4 + e < N) word[0] |= 1u << (8 * e);
    } else if (sig < K) {
        constexpr long LIMIT = SR == 2 ? 32639 : 8355711;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (j4 + e >= N) continue;
            const long cell = order ? (long)order[j4 + e] : j4 + e;       // the j-th cell of the contraction
            const double v2 = (2.0 * values[sig * ld + cell] - dg.centre2) * dg.scale;
            long q = (long)v2;
            if ((double)q != v2 || q > LIMIT || q < -LIMIT) {
                bad = true;
                q = 0;
            }
#pragma unroll
            for (int b = SR - 1; b >= 0; --b) {
                const long lo = ((q + 128) & 255) - 128;       // balanced digit in [-128, 127]
                q = (q - lo) >> 8;
                word[b] |= (uint32_t)(lo & 255) << (8 * e);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < SR; ++b)
        *reinterpret_cast<uint32_t *>(planes + ((long)b * Kpad + R) * Npad + j4) = word[b];
    if (bad) atomicOr(flag, 1);
}

// The same planes for rows that fit in shared memory (N <= ~28 000 cells): the contraction's
// cell order is a space-filling curve, so the kernel above reads a row of `values` 8 bytes at a
// time from scattered 32-byte sectors (4x the row in L2 -> SM traffic: 0.31 ms per projection at
// 3 500 x 20 000, the L2 rate).  Here a CTA copies the row into shared memory once (bulk
// asynchronous copies, row_tree.cuh) and permutes out of it; the next row is prefetched into L2.
constexpr int PACK_STAGED_THREADS = 1024;

template <int SR>
__global__ void __launch_bounds__(PACK_STAGED_THREADS)
pack_values_staged_kernel(const double *__restrict__ values, long K, long N, long ld, long Kpad, long Npad,
                          const int32_t *__restrict__ order, int8_t *__restrict__ planes,
                          const unsigned long long *__restrict__ range, int *__restrict__ flag) {
    extern __shared__ __align__(16) double row_s[];
    __shared__ uint64_t bar;
    const Digits dg = decode_range<SR>(range);
    const long n_bulk = N & ~1L;
    const uint32_t bytes = (uint32_t)(n_bulk * 8);
    // signature of tile row R, or -1 for the ones rows and the padding below the last signature
    auto sig_of = [&](long R) -> long {
        if (R >= Kpad) return -1;
        const long r = R % TM, sig = (R / TM) * SIGS_PER_TILE + r;
        return (r == TM - 1 || sig >= K) ? -1 : sig;
    };
    if (threadIdx.x == 0) {
        fp::rowtree::stage_init(&bar);
        const long sig = sig_of(blockIdx.x);
        if (sig >= 0 && n_bulk) fp::rowtree::stage_row_async(row_s, values + sig * ld, bytes, &bar);
    }
    __syncthreads();
    uint32_t parity = 0;
    bool bad = !dg.valid;
    for (long R = blockIdx.x; R < Kpad; R += gridDim.x) {
        const long sig = sig_of(R), next_sig = sig_of(R + gridDim.x), after = sig_of(R + 2 * (long)gridDim.x);
        if (threadIdx.x == 32 && after >= 0 && n_bulk) fp::rowtree::prefetch_row_l2(values + after * ld, bytes);
        if (sig >= 0) {
            if (n_bulk) {
                fp::rowtree::stage_wait(&bar, parity);
                parity ^= 1;
            }
            if (N & 1) {
                if (threadIdx.x == 0) row_s[N - 1] = values[sig * ld + N - 1];
                __syncthreads();
            }
        }
        const bool ones = (R % TM) == TM - 1;
        for (long j4 = (long)threadIdx.x * 4; j4 < Npad; j4 += PACK_STAGED_THREADS * 4) {
            uint32_t word[SR];
#pragma unroll
            for (int b = 0; b < SR; ++b) word[b] = 0;
            if (ones) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j4 + e < N) word[0] |= 1u << (8 * e);
            } else if (sig >= 0) {
                constexpr long LIMIT = SR == 2 ? 32639 : 8355711;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (j4 + e >= N) continue;
                    const long cell = order ? (long)order[j4 + e] : j4 + e;
                    const double v2 = (2.0 * row_s[cell] - dg.centre2) * dg.scale;
                    long q = (long)v2;
                    if ((double)q != v2 || q > LIMIT || q < -LIMIT) {
                        bad = true;
                        q = 0;
                    }
#pragma unroll
                    for (int b = SR - 1; b >= 0; --b) {
                        const long lo = ((q + 128) & 255) - 128;
                        q = (q - lo) >> 8;
                        word[b] |= (uint32_t)(lo & 255) << (8 * e);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < SR; ++b)
                *reinterpret_cast<uint32_t *>(planes + ((long)b * Kpad + R) * Npad + j4) = word[b];
        }
        __syncthreads();                                  // the row buffer is free again
        if (threadIdx.x == 0 && next_sig >= 0 && n_bulk)
            fp::rowtree::stage_row_async(row_s, values + next_sig * ld, bytes, &bar);
    }
    if (bad) atomicOr(flag, 1);
}

// Packed values (fp_values_pack) hold the digit planes in the CALLER's cell order; a projection
// needs them in its own order.  One CTA per plane row: the row (Npad bytes) is copied to shared
// memory, then written out permuted, four cells per store.  Rows too long for shared memory are
// gathered from global memory.
constexpr int PERMUTE_THREADS = 256;
constexpr long PERMUTE_SMEM_MAX = 200 * 1024;

__global__ void __launch_bounds__(PERMUTE_THREADS)
permute_planes_kernel(const int8_t *__restrict__ src, int8_t *__restrict__ dst, const int32_t *__restrict__ order,
                      long N, long Npad, int staged) {
    extern __shared__ __align__(16) int8_t prow[];
    const int8_t *row = src + (long)blockIdx.x * Npad;
    int8_t *out = dst + (long)blockIdx.x * Npad;
    if (staged) {
        for (long i = (long)threadIdx.x * 16; i < Npad; i += PERMUTE_THREADS * 16)
            *reinterpret_cast<uint4 *>(prow + i) = *reinterpret_cast<const uint4 *>(row + i);
        __syncthreads();
    }
    const int8_t *from = staged ? prow : row;
    for (long j4 = (long)threadIdx.x * 4; j4 < Npad; j4 += PERMUTE_THREADS * 4) {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (j4 + e < N) word |= (uint32_t)(uint8_t)from[order[j4 + e]] << (8 * e);
        *reinterpret_cast<uint32_t *>(out + j4) = word;
    }
}

// ---------------------------------------------------------------------------
// 2^t * (2^32 - 1), t <= 0, rounded to nearest: the fixed-point weight.
// t = n + k/128 + g (|g| <= 2^-8): table for 2^(k/128), cubic for 2^g - 1
// (remainder < 2^-40), exponent n added to the exponent field, + 2^52 rounds to
// an integer.  tab points at this lane's copy of the table (stride TAB_COPIES).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t weight_fixed32(double t, const char *__restrict__ tab_lane) {
    constexpr double MAGIC = 52776558133248.0;              // 1.5 * 2^45: ulp = 2^-7
    constexpr double L1 = 0.69314718055994530942, L2 = 0.24022650695910071233, L3 = 0.05550410866482157995;
    const double z = t + MAGIC;
    const int ti = __double2loint(z);                       // round(128 t), two's complement
    const double g = t - (z - MAGIC);                       // |g| <= 2^-8
    const double p = g * fma(g, fma(g, L3, L2), L1);        // 2^g - 1
    // 2^(k/128) * (2^32 - 1), k = ti mod 128; the lane's copy of entry k is 128 bytes after entry k-1
    const double T = *reinterpret_cast<const double *>(tab_lane + ((ti & (TAB - 1)) << 7));
    const double v = fma(T, p, T);
    // floor(t) <= 0 goes into the exponent field; below 2^-64 the weight rounds to 0 anyway, and the
    // clamp keeps the exponent arithmetic in range for any t > -2^24 (cells 1000 bandwidths apart)
    const int q = max(ti >> 7, -64);
    const double vs = __hiloint2double(__double2hiint(v) + (q << 20), __double2loint(v));
    return (uint32_t)__double2loint(vs + 4503599627370496.0);   // + 2^52: round to integer
}

// The MMAs of one K step (32 cells j): every (weight digit a, value digit b) pair with
// a + b < NCLASS, accumulated into class a + b.  Descriptors differ from the step's base
// descriptors by compile-time constants (plane offsets) and the accumulate flags are
// compile-time too, so the issuing thread executes little more than the MMAs themselves --
// with run-time flags and descriptor arithmetic it needed 336 instructions per stage and
// issued slower (93 clk per MMA) than the tensor core executes (64 clk).
// TZ: the TZ most significant weight digit planes of this (cell tile x stage) block are all zero
// (weight_planes_kernel's mask): their products are exactly zero and are not issued.
template <int SR, bool OPEN_CLASSES, int TZ>
__device__ __forceinline__ void issue_k_step(uint32_t tmem, uint64_t a_desc0, uint64_t b_desc0, uint32_t idesc) {
    static_assert(!OPEN_CLASSES || TZ == 0, "the first stage of a chunk initialises every class");
#pragma unroll
    for (int a = TZ; a < WD; ++a)
#pragma unroll
        for (int b = 0; b < SR; ++b) {
            const int c = a + b;                  // value digit 0 = most significant
            if (c >= NCLASS) continue;
            // first product written to a class: (0, c) for c < SR, else (c-SR+1, SR-1)
            const bool opens = (a == 0) || (b == SR - 1 && c >= SR);
            const uint64_t ad = a_desc0 + (uint64_t)(b * (A_PLANE_BYTES >> 4));
            const uint64_t bd = b_desc0 + (uint64_t)(a * (B_PLANE_BYTES >> 4));
            if (OPEN_CLASSES && opens)
                umma_i8_pair_c<0>(tmem + c * TN, ad, bd, idesc);
            else
                umma_i8_pair_c<1>(tmem + c * TN, ad, bd, idesc);
        }
}

template <int SR, int TZ>
__device__ __forceinline__ void issue_stage(uint32_t tmem, uint64_t a_st, uint64_t b_st, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < KS / 32; ++ks)          // K step ks: 32 bytes further along the swizzled rows
        issue_k_step<SR, false, TZ>(tmem, a_st + (uint64_t)(ks * 2), b_st + (uint64_t)(ks * 2), idesc);
}

// ---------------------------------------------------------------------------
// weight digit planes of a slab of cells i (all j): one thread owns 8 consecutive cells j
// (their records stay in registers) and walks 32 cells i, so a warp stores 256 contiguous
// bytes per digit plane and row.
// ---------------------------------------------------------------------------
constexpr int WG_THREADS = 256;
constexpr int WG_COLS = 8;                 // cells j per thread: one 8-byte store per digit plane
constexpr int WG_ROWS = 64;                // cells i per CTA

// planes[a][i - i_first][j] = digit a (0 = most significant) of W_ij; rows of Npad bytes
__global__ void __launch_bounds__(WG_THREADS, 3)
weight_planes_kernel(const CellI *__restrict__ ci, const CellJ *__restrict__ cj, const double *__restrict__ table,
                     long i_first, long slab_rows, long plane_rows, long Npad, uint8_t *__restrict__ planes,
                     uint32_t *__restrict__ block_or, int n_stages) {
    __shared__ __align__(16) double tab[TAB * TAB_COPIES];
    __shared__ CellI rows_s[WG_ROWS];
    for (int idx = threadIdx.x; idx < TAB * TAB_COPIES; idx += WG_THREADS) tab[idx] = table[idx];
    const long r0 = (long)blockIdx.y * WG_ROWS;                   // slab-relative first row
    if (threadIdx.x < WG_ROWS) rows_s[threadIdx.x] = ci[i_first + r0 + threadIdx.x];
    __syncthreads();
    const long j0 = ((long)blockIdx.x * WG_THREADS + threadIdx.x) * WG_COLS;
    if (j0 >= Npad) return;
    const char *my_tab = reinterpret_cast<const char *>(tab + (threadIdx.x & (TAB_COPIES - 1)));
    double jx[WG_COLS], jy[WG_COLS], jn[WG_COLS];
#pragma unroll
    for (int e = 0; e < WG_COLS; ++e) {
        const CellJ r = cj[j0 + e];
        jx[e] = r.x;
        jy[e] = r.y;
        jn[e] = r.nb;
    }
    const long rows = min((long)WG_ROWS, slab_rows - r0);
    uint32_t seen = 0;                        // OR of every weight this thread generates
#pragma unroll 1
    for (long r = 0; r < rows; ++r) {
        const CellI me = rows_s[r];
        uint32_t w[WG_COLS];
#pragma unroll
        for (int e = 0; e < WG_COLS; ++e) {
            double t = fma(me.c, jn[e], me.alpha);
            t = fma(me.x2c, jx[e], t);
            t = fma(me.y2c, jy[e], t);
            w[e] = weight_fixed32(t, my_tab);
        }
        // weights[diag] = 0 (Signatures.py:399)
        const long dj = i_first + r0 + r - j0;
        if (dj >= 0 && dj < WG_COLS) {
#pragma unroll
            for (int e = 0; e < WG_COLS; ++e)
                if (e == dj) w[e] = 0;
        }
#pragma unroll
        for (int e = 0; e < WG_COLS; ++e) seen |= w[e];
#pragma unroll
        for (int a = 0; a < WD; ++a) {
            const uint32_t sel = (3 - a) * 0x11 + 0x40;           // byte (3-a) of x, byte (3-a) of y
            uint2 v;
            v.x = __byte_perm(__byte_perm(w[0], w[1], sel), __byte_perm(w[2], w[3], sel), 0x5410);
            v.y = __byte_perm(__byte_perm(w[4], w[5], sel), __byte_perm(w[6], w[7], sel), 0x5410);
            *reinterpret_cast<uint2 *>(planes + ((long)a * plane_rows + r0 + r) * Npad + j0) = v;
        }
    }
    // which digit planes of the (cell tile, stage) block hold anything: byte 3 - a of the OR is plane a.
    // Sixteen consecutive threads cover one stage of 128 cells j (Npad is a multiple of 128, so a half
    // warp is in range as a whole).
    const unsigned active = __activemask();
    for (int o = 8; o; o >>= 1) seen |= __shfl_xor_sync(active, seen, o);
    if ((threadIdx.x & 15) == 0 && seen)
        atomicOr(&block_or[(r0 / TN) * n_stages + j0 / KS], seen);
}

template <int SR>
struct MmaSmem {
#ifdef FP_TC_TIMING_TWO_STAGES
    static constexpr int stages = 2;
#else
    static constexpr int stages = SR == 2 ? 3 : 2;
#endif
    static constexpr int stage_bytes = SR * A_PLANE_BYTES + WD * B_PLANE_BYTES;
    static constexpr int den_off = stages * stage_bytes;     // sum_j W_ij per column, then its scaled reciprocal
    static constexpr int rden_off = den_off + TN * 8;
    static constexpr int epi_off = rden_off + TN * 8;        // per epilogue warp: 32 rows x 8 columns of doubles
    static constexpr int bar_off = epi_off + EPI_WARPS * 2048;
    static constexpr int total = bar_off + 256;
};
static_assert(MmaSmem<2>::total <= 232448 && MmaSmem<3>::total <= 232448, "shared memory of the MMA kernel");

// Persistent pairs of CTAs.  A work item is 2 x 127 signatures (+ ones rows) x 128 cells i (one
// cell tile of the slab); value planes (A) and weight planes (B) both arrive by TMA.  Pair p
// takes the items p, p + pairs, p + 2 pairs, ... with the pair of signature blocks as the
// fast index: the pairs in flight share a cell tile, so its weight planes are read from HBM
// once and from L2 by the other signature blocks.  The barriers, tensor memory and descriptors
// are set up once per CTA, and the TMA producer runs on into the next item while the epilogue
// warps drain the accumulators of the current one.
template <int SR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
consistency_mma_kernel(const __grid_constant__ CUtensorMap tmap_values, const __grid_constant__ CUtensorMap tmap_weights,
                       const double *__restrict__ values, long ld, double *__restrict__ out, long ld_out, long N,
                       long K, long Kpad, long i_first, long plane_rows, int n_sig_pairs, int n_items,
                       int n_stages, int chunk_stages, int output_kind,
                       const unsigned long long *__restrict__ range, const int *__restrict__ flag,
                       const uint32_t *__restrict__ block_or, const int8_t *__restrict__ value_planes, long Npad,
                       const int32_t *__restrict__ out_order) {
    using L = MmaSmem<SR>;
    constexpr int STAGES = L::stages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    double *den_s = reinterpret_cast<double *>(smem + L::den_off);
    double *rden_s = reinterpret_cast<double *>(smem + L::rden_off);
    double *epi_s = reinterpret_cast<double *>(smem + L::epi_off);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::bar_off);          // [STAGES] leader's are used
    uint64_t *empty = full + 4;                                                  // [STAGES] one per CTA
    uint64_t *accum_full = empty + 4;                                            // one per CTA
    uint64_t *tmem_empty = accum_full + 1;                                       // leader's is used
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
    const int n_chunks = (n_stages + chunk_stages - 1) / chunk_stages;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 2);                           // one arrival per CTA's producer (the leader's carries the
                                                              // expect_tx); both CTAs' TMA bytes land on the leader's
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_full, 1);
        mbar_init(tmem_empty, 2 * EPI_WARPS);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_values);
        tma_prefetch_desc(&tmap_weights);
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
    if (tid >= 64 && tid - 64 < TN) den_s[tid - 64] = 0.0;
    tc_fence_before_sync();
    cluster_sync();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: this CTA's 128 value rows and
        // its 64 of the tile's 128 weight rows, all signalled on the leader's barrier
        if (lane == 0) {
            // both CTAs' bytes of a stage are counted on the leader's barrier: the leader cannot run
            // ahead of a peer that has not issued its loads, so neither producer is ever lapped
            // (Round 2 tried an L2 evict_last hint on the value planes, which every cell tile
            // re-reads: the 140 MB of planes then pushed the weight planes out before their 14
            // readers were through -- 13.1 GB of DRAM reads per launch instead of 9.5, same time:
            // the kernel is bound by the tensor pipe.  Plain loads stay.)
            long g = 0;                                       // stages issued by this CTA so far
            for (int item = pair; item < n_items; item += n_pairs) {
#ifdef FP_TC_TIMING_SAME_TILE
                // timing experiment (never a release build): every pair streams the SAME operands, so
                // every load hits L2 -- what the load path delivers without DRAM behind it
                const long tile_row = 0, blk = rank;
#else
                const long tile_row = (long)(item / n_sig_pairs) * TN;       // slab-relative first cell
                const long blk = (long)(item % n_sig_pairs) * 2 + rank;       // this CTA's block of 127 signatures
#endif
                // the issuer skips the products of all-zero leading weight planes (block_or): those
                // planes are not loaded either, and a stage with nothing to add loads nothing at all
                const uint32_t *seen_row = block_or + (long)(item / n_sig_pairs) * n_stages;
                uint32_t seen_next = __ldg(seen_row);
                for (int it = 0; it < n_stages; ++it, ++g) {
                    const int s = (int)(g % STAGES);
                    const uint32_t ph = (uint32_t)(g / STAGES) & 1;
                    const uint32_t seen = seen_next;
                    if (it + 1 < n_stages) seen_next = __ldg(seen_row + it + 1);
#ifdef FP_TC_TIMING_NO_SKIP
                    const int tz = 0;                         // timing experiment: every plane, no block_or
#else
                    const int tz = (it % chunk_stages == 0) ? 0 : (seen ? (__clz((int)seen) >> 3) : 4);
#endif
#ifdef FP_TC_SPIN_WAITS
                    mbar_spin(&empty[s], ph ^ 1);
#else
                    mbar_wait(&empty[s], ph ^ 1);
#endif
                    // both producers arrive on the leader's barrier for EVERY stage (also one that loads
                    // nothing), and both CTAs' bytes are counted there: the leader cannot run ahead of a
                    // peer that has not reached the stage, so neither producer is ever lapped
                    const uint32_t leader_full = mapa_shared(smem_u32(&full[s]), 0);
#ifdef FP_TC_TIMING_NO_LOADS
                    // timing experiment (never a release build): no operand traffic at all, the MMAs
                    // run on whatever shared memory holds -- what the issue path and the tensor pipe
                    // take by themselves
                    if (rank == 0)
                        mbar_arrive_expect_tx(&full[s], 0u);
                    else
                        mbar_arrive_cluster(leader_full);
                    continue;
#endif
                    if (rank == 0)
                        mbar_arrive_expect_tx(&full[s], tz == 4 ? 0u
                                                                : 2u * (uint32_t)(SR * A_PLANE_BYTES +
                                                                                  (WD - tz) * B_PLANE_BYTES));
                    else
                        mbar_arrive_cluster(leader_full);
                    if (tz == 4) continue;
                    uint8_t *dst = smem + s * L::stage_bytes;
#pragma unroll
                    for (int b = 0; b < SR; ++b)
                        tma_load_2d_pair(dst + b * A_PLANE_BYTES, &tmap_values, it * KS, (int)(b * Kpad + blk * TM),
                                         leader_full);
#pragma unroll
                    for (int a = 0; a < WD; ++a)
                        if (a >= tz)
                            tma_load_2d_pair(dst + SR * A_PLANE_BYTES + a * B_PLANE_BYTES, &tmap_weights, it * KS,
                                             (int)(a * plane_rows + tile_row + rank * TNH), leader_full);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (leader CTA)
        if (rank == 0) {
            const uint32_t idesc = idesc_i8(2 * TM, TN, /*a signed*/ 1, /*b unsigned*/ 0);
            const uint64_t a_desc_base = smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
            const uint64_t b_desc_base = smem_desc(smem_u32(smem) + SR * A_PLANE_BYTES, 16, 1024, LAYOUT_SW128);
            long g = 0, c = 0;                                // stages / accumulator chunks so far
            for (int item = pair; item < n_items; item += n_pairs) {
                // non-zero digit planes of this cell tile's blocks, one word per stage (read one stage ahead)
                const uint32_t *seen_row = block_or + (long)(item / n_sig_pairs) * n_stages;
                uint32_t seen_next = __ldg(seen_row);
                for (int it = 0; it < n_stages; ++it, ++g) {
                    const int s = (int)(g % STAGES);
                    const uint32_t ph = (uint32_t)(g / STAGES) & 1;
                    const bool chunk_first = it % chunk_stages == 0;
                    const bool chunk_last = it % chunk_stages == chunk_stages - 1 || it == n_stages - 1;
                    const uint32_t seen = seen_next;
                    if (it + 1 < n_stages) seen_next = __ldg(seen_row + it + 1);
                    // leading all-zero planes: their products are exactly zero (4: nothing to add at all)
#ifdef FP_TC_TIMING_NO_SKIP
                    const int tz = 0;
#else
                    const int tz = seen ? (__clz((int)seen) >> 3) : 4;
#endif
                    if (chunk_first && c > 0) {
                        mbar_wait_cluster(tmem_empty, (uint32_t)(c - 1) & 1);   // epilogues of both CTAs drained TMEM
                        tc_fence_after_sync();
                    }
#ifdef FP_TC_SPIN_WAITS
                    mbar_spin_cluster(&full[s], ph);
#else
                    mbar_wait_cluster(&full[s], ph);          // both CTAs' operands of this stage
#endif
                    tc_fence_after_sync();
                    if (elect_one()) {
                        const uint64_t a_st = a_desc_base + (uint64_t)(s * (L::stage_bytes >> 4));
                        const uint64_t b_st = b_desc_base + (uint64_t)(s * (L::stage_bytes >> 4));
#ifndef FP_TC_TIMING_NO_MMA
                        if (chunk_first) {
                            // the first stage of a chunk initialises every class: all products
                            issue_k_step<SR, true, 0>(tmem, a_st, b_st, idesc);
#pragma unroll
                            for (int ks = 1; ks < KS / 32; ++ks)
                                issue_k_step<SR, false, 0>(tmem, a_st + (uint64_t)(ks * 2), b_st + (uint64_t)(ks * 2),
                                                           idesc);
                        } else {
                            switch (tz) {
                                case 0: issue_stage<SR, 0>(tmem, a_st, b_st, idesc); break;
                                case 1: issue_stage<SR, 1>(tmem, a_st, b_st, idesc); break;
                                case 2: issue_stage<SR, 2>(tmem, a_st, b_st, idesc); break;
                                case 3: issue_stage<SR, 3>(tmem, a_st, b_st, idesc); break;
                                default: break;
                            }
                        }
#endif
#ifdef FP_TC_TIMING_PLAIN_ARRIVE
                        // timing experiment with -DFP_TC_TIMING_NO_MMA only (there is nothing to commit):
                        // the stage is handed back by plain arrivals instead of tcgen05.commit
                        mbar_arrive(&empty[s]);
                        mbar_arrive_cluster(mapa_shared(smem_u32(&empty[s]), 1));
#else
                        umma_commit_pair(&empty[s]);          // frees the stage in both CTAs
#endif
                        if (chunk_last) umma_commit_pair(accum_full);
                    }
                    __syncwarp();
                    if (chunk_last) ++c;
                }
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps (drain at the end of every j chunk)
        const int quad = warp & 3;                             // TMEM lanes 32*quad .. +31
        const int part = (warp - 2) >> 2;                      // columns 32*part .. +31 (4 warps per quadrant)
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        double *stage = epi_s + (warp - 2) * 256;
        const Digits dg = decode_range<SR>(range);
        // the class sums carry 256^(3-c); the exact product sum is 256^(SR-1) times larger
        const double centre = dg.centre2, inv_scale = (SR == 2 ? 256.0 : 65536.0) / dg.scale;
        const double unscale = 1.0 / dg.scale;                 // a power of two: exact
        const bool invalid = *flag != 0;
        long c = 0;                                            // accumulator chunks drained so far
#pragma unroll 1
        for (int item = pair; item < n_items; item += n_pairs) {
          const long i0 = i_first + (long)(item / n_sig_pairs) * TN;
          const long blk = (long)(item % n_sig_pairs) * 2 + rank;
#pragma unroll 1
          for (int chunk = 0; chunk < n_chunks; ++chunk, ++c) {
            const bool last = chunk == n_chunks - 1;
            mbar_wait(accum_full, (uint32_t)c & 1);
            tc_fence_after_sync();
            // phase A: the ones row (tile row 127) accumulates sum_j W_ij of every column
            if (quad == 3) {
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    const int col0 = part * 32 + b * 16;
                    uint32_t acc[NCLASS][16];
#pragma unroll
                    for (int cl = 0; cl < NCLASS; ++cl) tmem_ld16(lane_addr + cl * TN + col0, acc[cl]);
                    tmem_ld_wait();
                    if (lane == 31) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            long long d = 0;
#pragma unroll
                            for (int cl = 0; cl < NCLASS; ++cl)
                                d += (long long)(int)acc[cl][e] << (8 * (NCLASS - 1 - cl));
                            den_s[col0 + e] += (double)d;
                        }
                    }
                }
            }
            named_barrier_sync(1, EPI_THREADS);
            if (last) {
                // one division per column instead of one per output; zero row sum -> prediction 0
                // (Signatures.py:400-402)
                if (tid - 64 < TN) {
                    const double den = den_s[tid - 64];
                    rden_s[tid - 64] = den > 0.0 ? inv_scale / den : __longlong_as_double(0x7ff8000000000000LL);
                    den_s[tid - 64] = 0.0;                     // ready for the next item
                }
                named_barrier_sync(1, EPI_THREADS);
            }
            // Output.  A thread holds one signature row x 16 columns (TMEM lane = row), so writing
            // from that layout sends every store to 32 different rows (32 L1 wavefronts per
            // instruction; measured ~38 us per item, a sixth of the kernel).  The exact sums go
            // through a per-warp 32 x 8 staging tile (XOR-swizzled: two-way conflicts only, the
            // minimum for 8-byte words) and come back with 8 consecutive columns of 4 rows per
            // instruction: all global reads and writes are 64-byte row segments.
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const int col0 = part * 32 + b * 16;
                uint32_t acc[NCLASS][16];
#pragma unroll
                for (int cl = 0; cl < NCLASS; ++cl) tmem_ld16(lane_addr + cl * TN + col0, acc[cl]);
                tmem_ld_wait();
                if (b == 1) {
                    // this warp's last read of tensor memory: the next chunk (of this or the next item)
                    // may overwrite it once every epilogue warp of the pair has got this far -- the
                    // arithmetic and the global traffic below no longer hold the MMAs back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(tmem_empty), 0));
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        long long num = 0;
#pragma unroll
                        for (int cl = 0; cl < NCLASS; ++cl)
                            num += (long long)(int)acc[cl][8 * h + e] << (8 * (NCLASS - 1 - cl));
                        stage[lane * 8 + (e ^ ((lane >> 1) & 7))] = (double)num;
                    }
                    __syncwarp();
                    const int cc = lane & 7;
                    const long col = i0 + col0 + 8 * h + cc;
                    const double rden = rden_s[col0 + 8 * h + cc];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int rr = 4 * k + (lane >> 3);              // row within this warp's TMEM quadrant
                        double total = stage[rr * 8 + (cc ^ ((rr >> 1) & 7))];
                        const int trow = quad * 32 + rr;
                        const long tsig = blk * SIGS_PER_TILE + trow;
                        if (trow < SIGS_PER_TILE && tsig < K && col < N) {
                            // col is a position in the contraction's cell order; the output column is
                            // that cell's own index unless the caller asked for any order
                            double *optr = out + tsig * ld_out + (out_order ? (long)out_order[col] : col);
                            // partial sums of earlier j chunks are parked in the output row
                            if (chunk > 0) total += *optr;
                            if (!last) {
                                *optr = total;
                            } else {
                                // total = sum_j W_ij r_j, den = sum_j W_ij; v = (r / scale + centre2) / 2
                                double r = rden == rden ? 0.5 * (total * rden + centre) : 0.0;
                                if (output_kind == FP_OUT_ABS_ERROR) {
                                    // the cell's own value, rebuilt exactly from its digits (the value
                                    // planes are in the contraction's cell order; `values` is not)
                                    long q = 0;
#pragma unroll
                                    for (int b = 0; b < SR; ++b)
                                        q = q * 256 + (long)value_planes[((long)b * Kpad + blk * TM + trow) * Npad + col];
                                    r = fabs(0.5 * ((double)q * unscale + centre) - r);
                                }
                                if (invalid) r = __longlong_as_double(0x7ff8000000000000LL);
                                *optr = r;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // (rden_s is rewritten at the end of the next item at the earliest, two named barriers on)
          }
        }
    }
    tc_fence_before_sync();
    cluster_sync();
    if (warp == 0) tmem_dealloc_pair(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Plan {
    int sr;
    long nblk, Kpad, Npad;
    long slab_tiles;                             // cell tiles (128 cells i) whose weight planes are resident at once
    int nn_grid;                                 // cells per side of the nearest-neighbour grid (0: brute force)
    size_t nn_xy_off, nn_id_off, nn_cell_off, nn_cnt_off, nn_scan_off, nn_scan_bytes;
    size_t order_off, zkey_off, zsort_off, zsort_bytes;   // Z-order of the cells (only with the grid: N >= 4 096)
    size_t seen_off;                                      // non-zero digit planes per (cell tile, stage) of a slab
    size_t planes_off, ci_off, cj_off, nn_off, flag_off, range_off, table_off, weights_off, total;
};

constexpr size_t kWeightSlabBudget = 6ull << 30;   // bytes of generated weight planes kept at a time
constexpr long kGridNnMinCells = 4096;             // below this the brute-force nearest-neighbour kernel is as fast

// FASTPROJECT_B200_SLAB_BYTES overrides the budget (tests use it to run many slabs at small N)
size_t weight_slab_budget() {
    const char *e = getenv("FASTPROJECT_B200_SLAB_BYTES");
    if (e) {
        const long long v = atoll(e);
        if (v > 0) return (size_t)v;
    }
    return kWeightSlabBudget;
}

// Tiles per slab: as many as the budget holds, trimmed so that the CTA pairs of one slab fill
// whole waves of the machine (a slab is one launch; its last wave should not run half empty).
long pick_slab_tiles(long n_tiles, long sig_pairs, long Npad) {
    const size_t tile_bytes = (size_t)WD * TN * Npad;
    long cap = (long)(weight_slab_budget() / tile_bytes);
    if (cap < 1) cap = 1;
    if (cap >= n_tiles) return n_tiles;
    const long pairs_per_wave = fp::sm_count() / 2;
    long best = cap;
    double best_fill = 0.0;
    for (long t = cap; t >= 1 && t > cap / 2; --t) {
        const long items = t * sig_pairs, waves = (items + pairs_per_wave - 1) / pairs_per_wave;
        const double fill = (double)items / (double)(waves * pairs_per_wave);
        if (fill > best_fill + 1e-9) {
            best_fill = fill;
            best = t;
        }
    }
    return best;
}

bool make_plan(long N, long K, Plan *p) {
    if (N > 4000000) return false;               // three balanced digits hold |r * scale| <= 8 355 711
    p->sr = N <= 32640 ? 2 : 3;
    p->nblk = (K + SIGS_PER_TILE - 1) / SIGS_PER_TILE;
    p->nblk = (p->nblk + 1) / 2 * 2;             // CTA pairs own two blocks
    p->Kpad = p->nblk * TM;
    p->Npad = (long)align_up((size_t)N, KS);
    size_t off = 0;
    p->planes_off = off;
    off += align_up((size_t)p->sr * p->Kpad * p->Npad, 1024);
    p->ci_off = off;
    off += align_up((size_t)p->Npad * sizeof(CellI), 1024);
    p->cj_off = off;
    off += align_up((size_t)p->Npad * sizeof(CellJ), 1024);
    p->nn_off = off;
    off += align_up((size_t)p->Npad * 8, 1024);
    p->flag_off = off;
    p->range_off = off + 64;
    off += 1024;
    p->table_off = off;
    off += align_up((size_t)TAB * TAB_COPIES * 8, 1024);
    // nearest-neighbour grid: about one cell per grid cell; small projections keep the brute-force kernel
    p->nn_grid = 0;
    if (N >= kGridNnMinCells) {
        int G = (int)ceil(sqrt((double)N));
        if (G > 4096) G = 4096;
        p->nn_grid = G;
        const size_t cells = (size_t)G * G + 1;
        p->nn_xy_off = off;
        off += align_up((size_t)p->Npad * 16, 1024);
        p->nn_id_off = off;
        off += align_up((size_t)p->Npad * 4, 1024);
        p->nn_cell_off = off;
        off += align_up((size_t)p->Npad * 4, 1024);
        p->nn_cnt_off = off;                     // counts | offsets | fill, `cells` ints each
        off += align_up(3 * cells * 4, 1024);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int *)nullptr, (int *)nullptr, (int)cells);
        p->nn_scan_bytes = scan_bytes;
        p->nn_scan_off = off;
        off += align_up(scan_bytes, 1024);
        // Z-order: keys in / out and cell numbers in / out (4 x N words), the order itself, CUB's scratch
        p->zkey_off = off;
        off += align_up((size_t)4 * N * 4, 1024);
        p->order_off = off;
        off += align_up((size_t)p->Npad * 4, 1024);
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (const int32_t *)nullptr, (int32_t *)nullptr, (int)N, 0, 32);
        p->zsort_bytes = sort_bytes;
        p->zsort_off = off;
        off += align_up(sort_bytes, 1024);
    }
    p->slab_tiles = pick_slab_tiles(p->Npad / TN, p->nblk / 2, p->Npad);
    p->seen_off = off;
    off += align_up((size_t)p->slab_tiles * (p->Npad / KS) * 4, 1024);
    p->weights_off = off;
    off += align_up((size_t)WD * p->slab_tiles * TN * p->Npad, 1024);
    p->total = off;
    return true;
}

}  // namespace

namespace {
// CTA pairs of the persistent kernel that can be resident at once (a GPC with an odd number of
// free SMs leaves one unpaired): the grid must not exceed it, or a late pair would start its
// share of the items when the others are done.
template <int SR>
int resident_pairs() {
    static int cached = 0;
    if (cached) return cached;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)fp::sm_count());
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = MmaSmem<SR>::total;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, consistency_mma_kernel<SR>, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = fp::sm_count() / 2;
    }
    cached = n;
    return n;
}
}  // namespace

size_t fp_consistency_tc_workspace_bytes(int64_t N, int64_t K) {
    Plan p;
    if (N <= 0 || K <= 0 || !make_plan(N, K, &p)) return 0;
    return p.total;
}

int fp_consistency_tc_status(const void *workspace, int64_t N, int64_t K, cudaStream_t st) {
    Plan p;
    if (!make_plan(N, K, &p)) {
        fp::set_error("fp_consistency_status: FP_PRECISION_TC supports N <= 4000000 cells (N=%ld)", (long)N);
        return FP_EUNSUPPORTED;
    }
    int flag = 0;
    FP_CHECK_CUDA(cudaMemcpyAsync(&flag, static_cast<const char *>(workspace) + p.flag_off, sizeof(int),
                                  cudaMemcpyDeviceToHost, st));
    FP_CHECK_CUDA(cudaStreamSynchronize(st));
    if (flag) {
        fp::set_error("fp_consistency: FP_PRECISION_TC refused the values (NaN / inf, not half-integers, or a "
                      "range wider than three base-256 digits); every output of that call is NaN");
        return FP_EUNSUPPORTED;
    }
    return FP_OK;
}

// digit planes of `values` (through `order` when given), staged kernel for rows that fit in
// shared memory
int launch_pack_values(const Plan &p, const double *values, int64_t K, int64_t N, int64_t ld, const int32_t *order,
                       int8_t *planes, const unsigned long long *range, int *flag, cudaStream_t st) {
    const size_t staged_b = (size_t)((N + 1) & ~1L) * sizeof(double);
    if (fp::rowtree::stage_applies(values, ld, staged_b)) {
        static size_t staged_set[2] = {0, 0};
        auto kernel = p.sr == 2 ? pack_values_staged_kernel<2> : pack_values_staged_kernel<3>;
        if (staged_b > staged_set[p.sr == 3]) {
            FP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_b));
            staged_set[p.sr == 3] = staged_b;
        }
        long blocks = (long)fp::sm_count() * (staged_b <= 100 * 1024 ? 2 : 1);
        if (blocks > p.Kpad) blocks = p.Kpad;
        kernel<<<(unsigned)blocks, PACK_STAGED_THREADS, staged_b, st>>>(values, K, N, ld, p.Kpad, p.Npad, order, planes,
                                                                       range, flag);
        FP_CHECK_LAUNCH("pack_values_staged_kernel");
    } else {
        dim3 grid((unsigned)((p.Npad / 4 + 255) / 256), (unsigned)p.Kpad);
        if (p.sr == 2)
            pack_values_kernel<2><<<grid, 256, 0, st>>>(values, K, N, ld, p.Kpad, p.Npad, order, planes, range, flag);
        else
            pack_values_kernel<3><<<grid, 256, 0, st>>>(values, K, N, ld, p.Kpad, p.Npad, order, planes, range, flag);
        FP_CHECK_LAUNCH("pack_values_kernel");
    }
    return FP_OK;
}

// header of a packed-values blob (fp_values_pack): value range keys [3], then the refusal flag
constexpr size_t PACKED_FLAG_OFF = 32, PACKED_PLANES_OFF = 256;

size_t fp_values_pack_tc_bytes(int64_t N, int64_t K) {
    Plan p;
    if (N <= 0 || K <= 0 || !make_plan(N, K, &p)) return 0;
    return PACKED_PLANES_OFF + (size_t)p.sr * p.Kpad * p.Npad;
}

// value range + digit planes of `values` in the caller's cell order, into `packed`
int fp_values_pack_tc(const double *values, int64_t K, int64_t N, int64_t ld, void *packed, size_t packed_bytes,
                      cudaStream_t st) {
    Plan p;
    if (!make_plan(N, K, &p)) {
        fp::set_error("fp_values_pack: FP_PRECISION_TC supports N <= 4000000 cells (N=%ld)", (long)N);
        return FP_EUNSUPPORTED;
    }
    const size_t need = PACKED_PLANES_OFF + (size_t)p.sr * p.Kpad * p.Npad;
    if (!packed || packed_bytes < need) {
        fp::set_error("fp_values_pack: buffer %zu bytes < required %zu", packed_bytes, need);
        return FP_EWORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(packed) & 255) != 0) {
        fp::set_error("fp_values_pack: buffer must be 256-byte aligned");
        return FP_EINVAL;
    }
    char *pk = static_cast<char *>(packed);
    unsigned long long *range = reinterpret_cast<unsigned long long *>(pk);
    int *flag = reinterpret_cast<int *>(pk + PACKED_FLAG_OFF);
    int8_t *planes = reinterpret_cast<int8_t *>(pk + PACKED_PLANES_OFF);
    FP_CHECK_CUDA(cudaMemsetAsync(pk, 0, PACKED_PLANES_OFF, st));
    FP_CHECK_CUDA(cudaMemsetAsync(range, 0xff, 8, st));
    unsigned gy = (unsigned)(K < 296 ? K : 296);
    unsigned gx = (unsigned)((N + 2047) / 2048);
    if (gx > 4) gx = 4;
    value_range_kernel<<<dim3(gx, gy), 256, 0, st>>>(values, K, N, ld, range);
    FP_CHECK_LAUNCH("value_range_kernel");
    return launch_pack_values(p, values, K, N, ld, nullptr, planes, range, flag, st);
}

int fp_consistency_tc(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                      const double *values, const void *packed, int64_t N, int64_t K, int64_t ld, double *out,
                      int64_t ld_out, int output_kind, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    Plan p;
    if (!make_plan(N, K, &p)) {
        fp::set_error("fp_consistency: FP_PRECISION_TC supports N <= 4000000 cells (N=%ld)", (long)N);
        return FP_EUNSUPPORTED;
    }
    if (!workspace || workspace_bytes < p.total) {
        fp::set_error("fp_consistency: workspace %zu bytes < required %zu", workspace_bytes, p.total);
        return FP_EWORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) {
        fp::set_error("fp_consistency: workspace must be 1024-byte aligned");
        return FP_EINVAL;
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        fp::set_error("fp_consistency: cuTensorMapEncodeTiled not available from the driver");
        return FP_ECUDA;
    }
    char *ws = static_cast<char *>(workspace);
    int8_t *planes = reinterpret_cast<int8_t *>(ws + p.planes_off);
    CellI *ci = reinterpret_cast<CellI *>(ws + p.ci_off);
    CellJ *cj = reinterpret_cast<CellJ *>(ws + p.cj_off);
    int *flag = reinterpret_cast<int *>(ws + p.flag_off);

    unsigned long long *range = reinterpret_cast<unsigned long long *>(ws + p.range_off);
    unsigned long long *nn = reinterpret_cast<unsigned long long *>(ws + p.nn_off);
    const bool any_order = (output_kind & FP_OUT_ANY_CELL_ORDER) != 0;
    output_kind &= ~FP_OUT_ANY_CELL_ORDER;
    // FASTPROJECT_B200_NN=brute: the O(N^2) nearest-neighbour kernel; FASTPROJECT_B200_CELL_ORDER=given: no
    // Z-order (tests compare both with the defaults, bit for bit)
    const char *nn_env = getenv("FASTPROJECT_B200_NN");
    const bool brute = nn_env && !strcmp(nn_env, "brute");
    const char *order_env = getenv("FASTPROJECT_B200_CELL_ORDER");
    const bool use_grid = p.nn_grid && !brute;
    const bool z_order = use_grid && !(order_env && !strcmp(order_env, "given"));
    int32_t *order = z_order ? reinterpret_cast<int32_t *>(ws + p.order_off) : nullptr;
    {
        if (use_grid) {
            const int G = p.nn_grid;
            const size_t cells = (size_t)G * G + 1;
            double2 *sorted_xy = reinterpret_cast<double2 *>(ws + p.nn_xy_off);
            int *sorted_id = reinterpret_cast<int *>(ws + p.nn_id_off);
            int *cell_of = reinterpret_cast<int *>(ws + p.nn_cell_off);
            int *counts = reinterpret_cast<int *>(ws + p.nn_cnt_off);
            int *offsets = counts + cells, *fill = offsets + cells;
            unsigned long long *bounds = range + 4;           // behind the value range in the flag block
            FP_CHECK_CUDA(cudaMemsetAsync(bounds, 0xff, 16, st));
            FP_CHECK_CUDA(cudaMemsetAsync(bounds + 2, 0, 48, st));     // maxima and the four sums
            FP_CHECK_CUDA(cudaMemsetAsync(counts, 0, 3 * cells * 4, st));
            FP_CHECK_CUDA(cudaMemsetAsync(nn, 0x7f, (size_t)p.Npad * 8, st));   // padding rows: > +inf as a key
            const unsigned nb = (unsigned)((N + 255) / 256);
            nn_bounds_kernel<<<nb < 1024 ? nb : 1024, 256, 0, st>>>(coords, N, bounds);
            FP_CHECK_LAUNCH("nn_bounds_kernel");
            nn_count_kernel<<<nb, 256, 0, st>>>(coords, N, bounds, G, cell_of, counts);
            FP_CHECK_LAUNCH("nn_count_kernel");
            size_t scan_bytes = p.nn_scan_bytes;
            FP_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws + p.nn_scan_off, scan_bytes, counts, offsets, (int)cells, st));
            fp::count_launch(1);
            nn_scatter_kernel<<<nb, 256, 0, st>>>(coords, N, cell_of, offsets, fill, sorted_xy, sorted_id);
            FP_CHECK_LAUNCH("nn_scatter_kernel");
            nn_search_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(sorted_xy, sorted_id, N, offsets, bounds, G,
                                                                         nn);
            FP_CHECK_LAUNCH("nn_search_kernel");
            if (z_order) {
                // the cells along a Z-order curve over the same 3-sigma box: the contraction's cell order
                uint32_t *zk = reinterpret_cast<uint32_t *>(ws + p.zkey_off);
                uint32_t *zk_out = zk + N;
                int32_t *zc = reinterpret_cast<int32_t *>(zk_out + N);
                curve_key_kernel<<<nb, 256, 0, st>>>(coords, N, bounds, zk, zc);
                FP_CHECK_LAUNCH("curve_key_kernel");
                size_t sort_bytes = p.zsort_bytes;
                FP_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws + p.zsort_off, sort_bytes, zk, zk_out, zc, order, (int)N,
                                                              0, 32, st));
                fp::count_launch(4);
            }
        } else {
            FP_CHECK_CUDA(cudaMemsetAsync(nn, 0x7f, (size_t)p.Npad * 8, st));   // > +inf as a key
            const unsigned bx = (unsigned)((N + 255) / 256);
            unsigned by = (unsigned)((4 * fp::sm_count() + bx - 1) / bx);
            if (by < 1) by = 1;
            if ((long)by * 256 > N) by = (unsigned)((N + 255) / 256);
            nearest_neighbour_kernel<<<dim3(bx, by), 256, 0, st>>>(coords, N, nn);
            FP_CHECK_LAUNCH("nearest_neighbour_kernel");
        }
        cell_setup_kernel<<<(unsigned)((p.Npad + 255) / 256), 256, 0, st>>>(coords, sigma_sq, sigma_sq_scalar, N,
                                                                            p.Npad, nn, order, ci, cj,
                                                                            reinterpret_cast<double *>(ws + p.table_off));
        FP_CHECK_LAUNCH("cell_setup_kernel");
    }
    {
        // value range and digit planes (in the contraction's cell order): re-derived from `values` on
        // every call.  (Round 1 had a caller-asserted "same values as the previous call" flag that
        // skipped this; a wrong assertion gave silently wrong results, and verifying it costs as much
        // as packing.)
        FP_CHECK_CUDA(cudaMemsetAsync(flag, 0, 64, st));
        if (packed) {
            // the values arrive as fp_values_pack left them: range and refusal flag in the header, digit
            // planes in the caller's cell order -- only the permutation to this projection's order is left
            const char *pk = static_cast<const char *>(packed);
            FP_CHECK_CUDA(cudaMemcpyAsync(range, pk, 24, cudaMemcpyDeviceToDevice, st));
            FP_CHECK_CUDA(cudaMemcpyAsync(flag, pk + PACKED_FLAG_OFF, sizeof(int), cudaMemcpyDeviceToDevice, st));
            const int8_t *planes0 = reinterpret_cast<const int8_t *>(pk + PACKED_PLANES_OFF);
            const size_t plane_bytes = (size_t)p.sr * p.Kpad * p.Npad;
            if (!order) {
                FP_CHECK_CUDA(cudaMemcpyAsync(planes, planes0, plane_bytes, cudaMemcpyDeviceToDevice, st));
            } else {
                const int staged = p.Npad <= PERMUTE_SMEM_MAX;
                const size_t smem = staged ? (size_t)p.Npad : 0;
                static size_t smem_set = 0;
                if (smem > 40 * 1024 && smem > smem_set) {
                    FP_CHECK_CUDA(cudaFuncSetAttribute(permute_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)smem));
                    smem_set = smem;
                }
                permute_planes_kernel<<<(unsigned)(p.sr * p.Kpad), PERMUTE_THREADS, smem, st>>>(planes0, planes, order, N,
                                                                                             p.Npad, staged);
                FP_CHECK_LAUNCH("permute_planes_kernel");
            }
        } else {
            FP_CHECK_CUDA(cudaMemsetAsync(range, 0xff, 8, st));
            FP_CHECK_CUDA(cudaMemsetAsync(range + 1, 0, 16, st));
            unsigned gy = (unsigned)(K < 296 ? K : 296);
            unsigned gx = (unsigned)((N + 2047) / 2048);
            if (gx > 4) gx = 4;
            value_range_kernel<<<dim3(gx, gy), 256, 0, st>>>(values, K, N, ld, range);
            FP_CHECK_LAUNCH("value_range_kernel");
            const int rc = launch_pack_values(p, values, K, N, ld, order, planes, range, flag, st);
            if (rc != FP_OK) return rc;
        }
    }
    CUtensorMap tmap, tmap_w;
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.Npad, (cuuint64_t)(p.sr * p.Kpad)};
        cuuint64_t strides[1] = {(cuuint64_t)p.Npad};
        cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)TM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, planes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         FP_TC_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fp::set_error("fp_consistency: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return FP_ECUDA;
        }
    }
    const int n_stages = (int)(p.Npad / KS);
    static bool attr_set = false;
    if (!attr_set) {
        FP_CHECK_CUDA(cudaFuncSetAttribute(consistency_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           MmaSmem<2>::total));
        FP_CHECK_CUDA(cudaFuncSetAttribute(consistency_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           MmaSmem<3>::total));
        attr_set = true;
    }
    uint8_t *wplanes = reinterpret_cast<uint8_t *>(ws + p.weights_off);
    const long plane_rows = p.slab_tiles * TN;
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.Npad, (cuuint64_t)(WD * plane_rows)};
        cuuint64_t strides[1] = {(cuuint64_t)p.Npad};
        cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)TNH};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, wplanes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         FP_TC_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fp::set_error("fp_consistency: cuTensorMapEncodeTiled (weights) failed (%d)", (int)r);
            return FP_ECUDA;
        }
    }
    const long n_tiles = p.Npad / TN;
    for (long t0 = 0; t0 < n_tiles; t0 += p.slab_tiles) {
        const long tiles = n_tiles - t0 < p.slab_tiles ? n_tiles - t0 : p.slab_tiles;
        const long i_first = t0 * TN, slab_rows = tiles * TN;
        dim3 ggrid((unsigned)((p.Npad / WG_COLS + WG_THREADS - 1) / WG_THREADS), (unsigned)(slab_rows / WG_ROWS));
        uint32_t *seen = reinterpret_cast<uint32_t *>(ws + p.seen_off);
        FP_CHECK_CUDA(cudaMemsetAsync(seen, 0, (size_t)tiles * n_stages * 4, st));
        weight_planes_kernel<<<ggrid, WG_THREADS, 0, st>>>(ci, cj, reinterpret_cast<const double *>(ws + p.table_off),
                                                           i_first, slab_rows, plane_rows, p.Npad, wplanes, seen,
                                                           n_stages);
        FP_CHECK_LAUNCH("weight_planes_kernel");
        // persistent CTA pairs: one per pair of SMs, or fewer when the slab has fewer items
        const int n_sig_pairs = (int)(p.nblk / 2);
        const long n_items = tiles * n_sig_pairs;
        long n_pairs = p.sr == 2 ? resident_pairs<2>() : resident_pairs<3>();
        if (n_pairs > n_items) n_pairs = n_items;
        dim3 grid((unsigned)(2 * n_pairs));
        if (p.sr == 2)
            consistency_mma_kernel<2><<<grid, THREADS, MmaSmem<2>::total, st>>>(
                tmap, tmap_w, values, ld, out, ld_out, N, K, p.Kpad, i_first, plane_rows, n_sig_pairs, (int)n_items,
                n_stages, n_stages, output_kind, range, flag, seen, planes, p.Npad, any_order ? nullptr : order);
        else
            consistency_mma_kernel<3><<<grid, THREADS, MmaSmem<3>::total, st>>>(
                tmap, tmap_w, values, ld, out, ld_out, N, K, p.Kpad, i_first, plane_rows, n_sig_pairs, (int)n_items,
                n_stages, CHUNK_STAGES_3, output_kind, range, flag, seen, planes, p.Npad,
                any_order ? nullptr : order);
        FP_CHECK_LAUNCH("consistency_mma_kernel");
    }
    return FP_OK;
}
