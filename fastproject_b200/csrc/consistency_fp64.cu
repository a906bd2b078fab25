// Signature x projection local dissimilarity, float64 arithmetic end to end.
//
// Replaces, for one projection, /root/reference/FastProject/Signatures.py:396-408
// and :412-413: cdist -> exp(-d^2/sigma^2) -> zero diagonal -> row normalise ->
// np.dot(weights, rank matrix) -> |rank - prediction|, for real and random
// signatures together, WITHOUT materialising the N x N kernel matrix: every
// CTA owns a 128-cell x 128-signature output tile and streams the cells j in
// chunks of 16, recomputing the Gaussian weights of its 128 rows on the fly
// (flash-attention style: weights live only in shared memory).
//
// This is FP_PRECISION_FP64: the verifier / parity-grade mode.  The contraction
// runs as DFMA on the FP64 pipe with an 8x8 register tile per thread.  The
// tensor-core (tcgen05) mode lives in consistency_tc.cu.
#include "fp_common.cuh"

namespace {

constexpr int BM = 128;      // cells (rows i) per CTA
constexpr int BN = 128;      // signatures (columns k) per CTA
constexpr int BK = 16;       // cells j per chunk
constexpr int THREADS = 256; // 16 x 16 threads, 8 x 8 outputs each
constexpr int RS_PAD = 2;    // padding (doubles) of the rank tile rows

struct __align__(16) Smem {
    double w[BK][BM];              // kernel weights  w[j][i]
    double r[BK][BN + RS_PAD];     // ranks           r[j][k]
    double xj[BK], yj[BK];
    double rowsum[2][BM];
};

__global__ void __launch_bounds__(THREADS)
consistency_fp64_kernel(const double *__restrict__ coords, const double *__restrict__ sigma_sq,
                        double sigma_sq_scalar, const double *__restrict__ ranks, long N, long K,
                        long ld, double *__restrict__ out, long ld_out, int output_kind) {
    __shared__ Smem sm;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long i0 = (long)blockIdx.x * BM;
    const long k0 = (long)blockIdx.y * BN;
    const double *cx = coords, *cy = coords + N;

    // weight producer role: thread -> (row wi, eight j's of the chunk)
    const int wi = tid & (BM - 1);
    const int wj0 = (tid >> 7) * 8;
    const long gi = i0 + wi;
    const bool row_live = gi < N;
    const double xi = row_live ? cx[gi] : 0.0;
    const double yi = row_live ? cy[gi] : 0.0;
    const double s2 = sigma_sq ? (row_live ? sigma_sq[gi] : 1.0) : sigma_sq_scalar;
    const double neg_inv_s2 = -1.0 / s2;
    double my_rowsum = 0.0;

    // rank loader role: thread -> (signature lk, half of the 16 j's)
    const int lk = tid & (BN - 1);
    const int lj0 = (tid >> 7) * 8;
    const bool col_live = (k0 + lk) < K;
    const double *rrow = ranks + (k0 + (col_live ? lk : 0)) * ld;

    double acc[8][8];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[m][n] = 0.0;

    for (long j0 = 0; j0 < N; j0 += BK) {
        if (tid < BK) {
            const long gj = j0 + tid;
            sm.xj[tid] = gj < N ? cx[gj] : 0.0;
            sm.yj[tid] = gj < N ? cy[gj] : 0.0;
        }
        // ranks tile: r[j][k] <- ranks[k][j]
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const long gj = j0 + lj0 + jj;
            sm.r[lj0 + jj][lk] = (col_live && gj < N) ? rrow[gj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const long gj = j0 + wj0 + jj;
            const double dx = xi - sm.xj[wj0 + jj], dy = yi - sm.yj[wj0 + jj];
            const double d2 = dx * dx + dy * dy;
            double w = exp(d2 * neg_inv_s2);
            if (gj >= N || gj == gi || !row_live) w = 0.0;   // weights[diag] = 0 (:399)
            sm.w[wj0 + jj][wi] = w;
            my_rowsum += w;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < BK; ++jj) {
            double a[8], b[8];
            const double2 *wp = reinterpret_cast<const double2 *>(&sm.w[jj][ty * 8]);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double2 v = wp[m];
                a[2 * m] = v.x;
                a[2 * m + 1] = v.y;
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const double2 v = *reinterpret_cast<const double2 *>(&sm.r[jj][tx * 2 + 32 * n]);
                b[2 * n] = v.x;
                b[2 * n + 1] = v.y;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m)
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[m][n] = fma(a[m], b[n], acc[m][n]);
        }
        __syncthreads();
    }

    sm.rowsum[tid >> 7][wi] = my_rowsum;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int li = ty * 8 + m;
        const long gi2 = i0 + li;
        if (gi2 >= N) continue;
        double rs = sm.rowsum[0][li] + sm.rowsum[1][li];
        if (rs == 0.0) rs = 1.0;                              // weights_norm_factor == 0 -> 1 (:401)
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const long gk = k0 + tx * 2 + 32 * (n >> 1) + (n & 1);
            if (gk >= K) continue;
            const double pred = acc[m][n] / rs;
            out[gk * ld_out + gi2] = output_kind == FP_OUT_PREDICTION
                                         ? pred
                                         : fabs(ranks[gk * ld + gi2] - pred);   // (:408)
        }
    }
}

}  // namespace

int fp_consistency_fp64(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                        const double *ranks, int64_t N, int64_t K, int64_t ld, double *out_dissim,
                        int64_t ld_out, int output_kind, cudaStream_t st) {
    dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)((K + BN - 1) / BN));
    consistency_fp64_kernel<<<grid, THREADS, 0, st>>>(coords, sigma_sq, sigma_sq_scalar, ranks, N, K,
                                                      ld, out_dissim, ld_out, output_kind);
    FP_CHECK_LAUNCH("consistency_fp64_kernel");
    return FP_OK;
}
