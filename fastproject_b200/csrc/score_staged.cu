// Signature scoring, TMA-staged: the expression / weight / zero-mask matrices are read in
// (256 genes x 16 cells) boxes by cp.async.bulk.tensor into shared memory, and every signature of
// the CTA's group gathers its genes from there.
//
// Same arithmetic and the same order as score_gather_kernel (score.cu): per signature the genes
// in ascending order, every product a separately rounded __dmul_rn and every accumulation a
// __dadd_rn -- bit-identical to NumPy's evaluation of
// /root/reference/FastProject/SigScoreMethods.py:18-162.  Only the data path differs.
//
// Why.  A gene row is used by nnz/G signatures (15 at BASELINE config 2, 58 with the default
// 18 000 random signatures).  The gather kernel pulls it through L2 -> SM once per use: 16 * nnz
// * N bytes (74.6 GB at config 2) against the L2 throughput cap of the part (~6 300 B/clk, about
// 12 TB/s: LDG and TMA alike) -- 5.7 - 7.2 ms.  Here a CTA owns a tile of 16 cells and a GROUP of
// up to 512 signatures whose accumulators stay in registers for the whole pass over the genes
// (1 024 threads x 8 signatures); it streams all G gene rows of its tile once, so the L2 -> SM
// traffic is 16 * G * N * ceil(K / 512) bytes (33.6 GB at config 2) and the gathers hit shared
// memory: rows of 16 cells are 128 bytes = all 32 banks, a warp reads two rows per request
// (lanes 0-15 one signature, lanes 16-31 another), conflict-free by construction.
//
// Bounds: max(16 G N ceil(K/512) / L2 cap, 16 nnz N / shared-memory bandwidth (148 SMs x 128 B/clk),
// 16 G N / HBM).  Signature groups of one cell tile are adjacent in the grid, so they run
// concurrently and all but the first read a gene block from L2; the compulsory HBM traffic stays
// 16 G N + 8 N K.
//
// Pipeline (per CTA): 3 stages (2 with three operands) of one (256 x 16) box per operand, full /
// empty mbarriers; lane 0 of warp 0 is the producer -- after finishing gene block b it waits until
// all 32 warps have released that stage and issues the loads of block b + stages.  Signatures are
// dealt to the threads chunk-cyclically (chunk = 64 consecutive signatures, one per (warp, half
// warp); a thread's eight signatures are 64 * groups apart) so that every warp and every group gets
// the same mix of signature sizes.
#include <cuda.h>

#include "fp_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace fp::tc;

constexpr int ST_CELLS = 16;            // cells per CTA tile: 128-byte rows
constexpr int ST_GENES = 256;           // genes per box (TMA box dimensions are <= 256)
constexpr int ST_WARPS = 32;
constexpr int ST_THREADS = ST_WARPS * 32;
constexpr int ST_A = 8;                 // signatures per thread
constexpr int ST_CHUNK = ST_WARPS * 2;  // 64: one signature per (warp, half warp)
constexpr int ST_GROUP = ST_CHUNK * ST_A;   // 512 signatures per CTA
constexpr int ST_BOX_BYTES = ST_GENES * ST_CELLS * 8;   // 32 KB

template <int METHOD>
struct StagedCfg {
    static constexpr bool use_w = METHOD == FP_METHOD_WEIGHTED || METHOD == FP_METHOD_IMPUTED;
    static constexpr bool use_z = METHOD == FP_METHOD_NONZERO || METHOD == FP_METHOD_IMPUTED;
    static constexpr int n_ops = 1 + (use_w ? 1 : 0) + (use_z ? 1 : 0);
    static constexpr int stages = n_ops == 3 ? 2 : 3;
    static constexpr int stage_bytes = n_ops * ST_BOX_BYTES;
    static constexpr int smem_bytes = stages * stage_bytes + 128;
};

// ---------------------------------------------------------------------------------------------
// Preparation (one thread per signature, once per call): the CSR entries re-expressed for the
// staged kernel -- op[t] = (gene mod 256) | (sign < 0) << 8, and cut[sig][b] = first entry of
// signature sig whose gene lies in gene block >= b (b = 0 .. n_blocks).  With them the hot loop
// needs ONE 2-byte load per entry and ONE 4-byte load per (signature, gene block), none of them
// dependent on a previous entry: the first version walked the raw CSR with a dependent
// gene-index load per entry and per block visit and was latency bound at 18 ms (config 2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
score_prepare_kernel(const int32_t *__restrict__ sig_ptr, const int32_t *__restrict__ sig_gene,
                     const int8_t *__restrict__ sig_sign, long K, int n_blocks, uint16_t *__restrict__ op,
                     int32_t *__restrict__ cut) {
    // rows K .. rows_padded-1 (thread slots without a signature) stay all zero: empty ranges
    const long sig = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sig >= K) return;
    const int beg = sig_ptr[sig], end = sig_ptr[sig + 1];
    int32_t *row = cut + sig * (n_blocks + 1);
    int b = 0;
    row[0] = beg;
    for (int t = beg; t < end; ++t) {
        const int g = sig_gene[t];
        const int gb = g / ST_GENES;
        while (b < gb && b < n_blocks) row[++b] = t;
        op[t] = (uint16_t)((g - gb * ST_GENES) | (sig_sign[t] < 0 ? 0x100 : 0));
    }
    while (b < n_blocks) row[++b] = end;
}

template <int METHOD>
__global__ void __launch_bounds__(ST_THREADS, 1)
score_staged_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_z, long G, long N,
                    const int32_t *__restrict__ sig_ptr, const uint16_t *__restrict__ op,
                    const int32_t *__restrict__ cut, long K, int n_groups,
                    const double *__restrict__ gene_mu, double *__restrict__ out, long ld_out) {
    using C = StagedCfg<METHOD>;
    extern __shared__ __align__(128) unsigned char st_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(st_smem + C::stages * C::stage_bytes);
    uint64_t *empty = full + C::stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = lane >> 4, c = lane & 15;
    const long bid = blockIdx.x;
    const int group = (int)(bid % n_groups);
    const long cell0 = (bid / n_groups) * ST_CELLS;
    const int n_blocks = (int)((G + ST_GENES - 1) / ST_GENES);

    if (tid == 0) {
        for (int s = 0; s < C::stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ST_WARPS);
        }
        mbar_fence_init();
        tma_prefetch_desc(&map_x);
        if (C::use_w) tma_prefetch_desc(&map_w);
        if (C::use_z) tma_prefetch_desc(&map_z);
    }
    __syncthreads();

    auto issue = [&](int b) {          // producer thread only
        const int s = b % C::stages;
        unsigned char *dst = st_smem + s * C::stage_bytes;
        mbar_arrive_expect_tx(&full[s], (uint32_t)C::stage_bytes);
        tma_load_2d(dst, &map_x, (int)cell0, b * ST_GENES, &full[s]);
        int o = 1;
        if (C::use_w) tma_load_2d(dst + (o++) * ST_BOX_BYTES, &map_w, (int)cell0, b * ST_GENES, &full[s]);
        if (C::use_z) tma_load_2d(dst + o * ST_BOX_BYTES, &map_z, (int)cell0, b * ST_GENES, &full[s]);
    };
    if (tid == 0)
        for (int b = 0; b < C::stages && b < n_blocks; ++b) issue(b);

    // this thread's signatures: chunk (a * n_groups + group), slot (warp, half)
    // (the cut table has a zero row for every thread slot beyond K: its ranges are empty)
    int cur[ST_A];
    double num[ST_A], acc2[ST_A];      // acc2: the normaliser (weighted, non-zero) or the imputed part
    const int cut_ld = n_blocks + 1;
    const int32_t *cut0 = cut + ((long)group * ST_CHUNK + warp * 2 + half) * cut_ld;
    const long cut_step = (long)n_groups * ST_CHUNK * cut_ld;       // from signature slot a to a + 1
#pragma unroll
    for (int a = 0; a < ST_A; ++a) {
        cur[a] = __ldg(cut0 + a * cut_step);
        num[a] = 0.0;
        acc2[a] = 0.0;
    }

    for (int b = 0; b < n_blocks; ++b) {
        const int s = b % C::stages;
        const uint32_t ph = (uint32_t)(b / C::stages) & 1;
        // where each of my signatures stops in this gene block: eight independent loads, issued
        // before the wait for the block's data
        int stop[ST_A];
#pragma unroll
        for (int a = 0; a < ST_A; ++a) stop[a] = __ldg(cut0 + a * cut_step + b + 1);
        mbar_wait(&full[s], ph);
        const double *sx = reinterpret_cast<const double *>(st_smem + s * C::stage_bytes) + c;
        const double *sw = sx + (C::use_w ? ST_GENES * ST_CELLS : 0);
        const double *sz = sx + (C::use_z ? (C::use_w ? 2 : 1) * ST_GENES * ST_CELLS : 0);
        const int g_lo = b * ST_GENES;
#pragma unroll
        for (int a = 0; a < ST_A; ++a) {
            for (int t = cur[a]; t < stop[a]; ++t) {
                const unsigned o = __ldg(op + t);
                const int r = (int)(o & 0xff) * ST_CELLS;
                // x * (+-1): flip the sign bit (exactly the product NumPy rounds to)
                const double x = sx[r];
                const double xs = __hiloint2double(__double2hiint(x) ^ (int)((o & 0x100) << 23), __double2loint(x));
                if (METHOD == FP_METHOD_NAIVE) {
                    // pdata = data * sig * ones  (:38-41); the normaliser is the gene count
                    num[a] = __dadd_rn(num[a], xs);
                } else if (METHOD == FP_METHOD_WEIGHTED) {
                    const double w = sw[r];               // pdata = data * sig * weights  (:70-73)
                    num[a] = __dadd_rn(num[a], __dmul_rn(xs, w));
                    acc2[a] = __dadd_rn(acc2[a], w);
                } else if (METHOD == FP_METHOD_NONZERO) {
                    const double nz = __dsub_rn(1.0, sz[r]);      // (data * not_zeros) * sig  (:152-153)
                    const double p = __dmul_rn(x, nz);
                    num[a] = __dadd_rn(num[a], __hiloint2double(__double2hiint(p) ^ (int)((o & 0x100) << 23),
                                                                __double2loint(p)));
                    acc2[a] = __dadd_rn(acc2[a], nz);
                } else {                                           // imputed (:105-121)
                    const double w = sw[r], z = sz[r], mu = __ldg(gene_mu + g_lo + (int)(o & 0xff));
                    const double sgn = (o & 0x100) ? -1.0 : 1.0;
                    const double nz = __dsub_rn(1.0, z);
                    num[a] = __dadd_rn(num[a], __dmul_rn(__dmul_rn(x, nz), sgn));
                    const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, w), mu), sgn), z);
                    const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(w, x), sgn), z);
                    acc2[a] = __dadd_rn(acc2[a], __dadd_rn(t1, t2));
                }
            }
            cur[a] = stop[a];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (tid == 0 && b + C::stages < n_blocks) {
            mbar_wait(&empty[s], ph);           // every warp is done reading block b
            issue(b + C::stages);
        }
        __syncwarp();
    }

    const long cell = cell0 + c;
#pragma unroll
    for (int a = 0; a < ST_A; ++a) {
        const long sig = ((long)a * n_groups + group) * ST_CHUNK + warp * 2 + half;
        if (sig >= K || cell >= N) continue;
        const int n_k = sig_ptr[sig + 1] - sig_ptr[sig];
        double score;
        if (METHOD == FP_METHOD_NAIVE) {
            score = __ddiv_rn(num[a], (double)n_k);
        } else if (METHOD == FP_METHOD_IMPUTED) {
            score = __ddiv_rn(__dadd_rn(num[a], acc2[a]), (double)n_k);
        } else {
            double den = acc2[a];
            if (den == 0.0) den = 1.0;                    // norm_factor[norm_factor == 0] = 1.0
            score = __ddiv_rn(num[a], den);
        }
        out[sig * ld_out + cell] = score;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

bool encode_matrix(EncodeTiledFn enc, CUtensorMap *map, const double *base, long G, long N, long ld) {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)G};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)ST_CELLS, (cuuint32_t)ST_GENES};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int METHOD>
int launch_staged(const CUtensorMap &mx, const CUtensorMap &mw, const CUtensorMap &mz, long G, long N,
                  const int32_t *sig_ptr, const uint16_t *op, const int32_t *cut, long K, int n_groups,
                  const double *gene_mu, double *out, long ld_out, cudaStream_t st) {
    using C = StagedCfg<METHOD>;
    static bool attr_set = false;
    if (!attr_set) {
        FP_CHECK_CUDA(cudaFuncSetAttribute(score_staged_kernel<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           C::smem_bytes));
        attr_set = true;
    }
    const long tiles = (N + ST_CELLS - 1) / ST_CELLS;
    const long blocks = tiles * n_groups;
    if (blocks >= (1L << 31)) {
        fp::set_error("fp_score_signatures: grid of %ld CTAs", blocks);
        return FP_EUNSUPPORTED;
    }
    score_staged_kernel<METHOD><<<(unsigned)blocks, ST_THREADS, C::smem_bytes, st>>>(
        mx, mw, mz, G, N, sig_ptr, op, cut, K, n_groups, gene_mu, out, ld_out);
    FP_CHECK_LAUNCH("score_staged_kernel");
    return FP_OK;
}

struct StagedPlan {
    long n_groups, rows, n_blocks;
    size_t op_off, cut_off, total;
};

StagedPlan staged_plan(long G, long K, long nnz) {
    StagedPlan p;
    p.n_groups = (K + ST_GROUP - 1) / ST_GROUP;
    p.rows = p.n_groups * ST_GROUP;
    p.n_blocks = (G + ST_GENES - 1) / ST_GENES;
    p.op_off = 0;
    p.cut_off = ((size_t)nnz * 2 + 255) & ~(size_t)255;
    p.total = p.cut_off + (((size_t)p.rows * (p.n_blocks + 1) * 4 + 255) & ~(size_t)255);
    return p;
}

}  // namespace

size_t fp_score_staged_workspace_bytes(long G, long K, long nnz) {
    if (G <= 0 || K <= 0 || nnz <= 0) return 0;
    return staged_plan(G, K, nnz).total;
}

// Whether the staged kernel applies and pays: TMA needs 16-byte aligned bases and row pitches;
// it moves 16 G N ceil(K/512) bytes through L2 where the gather kernel moves 16 nnz N.
bool fp_score_staged_applies(const double *X, const double *W, const double *Z, long G, long N, long ld, long K,
                             long nnz, bool forced) {
    if (nnz <= 0 || K <= 0) return false;
    if ((ld & 1) || N < ST_CELLS || G >= (1L << 31) || N >= (1L << 31)) return false;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) |
                           reinterpret_cast<uintptr_t>(Z);
    if (bits & 15) return false;
    const StagedPlan p = staged_plan(G, K, nnz);
    if ((size_t)p.rows * (p.n_blocks + 1) >= (1ull << 31)) return false;
    // Measured on B200 (profiles/r02_scoring_staged.md): at BASELINE config 2 the staged kernel
    // takes 9.2 ms against 5.7 ms for the gathers -- the membership tables are re-read through
    // L2 (46 GB of sector traffic) and the 2 x 16-lane layout issues 3x the instructions per
    // gene -- so it is used only when asked for (FASTPROJECT_B200_SCORE=staged).
    (void)p;
    return forced;
}

int fp_score_staged(int method, const double *X, const double *W, const double *Z, long G, long N, long ld,
                    const int32_t *sig_ptr, const int32_t *sig_gene, const int8_t *sig_sign, long K, long nnz,
                    const double *gene_mu, double *out, long ld_out, void *workspace, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        fp::set_error("fp_score_signatures: cuTensorMapEncodeTiled not available from the driver");
        return FP_ECUDA;
    }
    CUtensorMap mx, mw, mz;
    const bool use_w = method == FP_METHOD_WEIGHTED || method == FP_METHOD_IMPUTED;
    const bool use_z = method == FP_METHOD_NONZERO || method == FP_METHOD_IMPUTED;
    if (!encode_matrix(enc, &mx, X, G, N, ld) || !encode_matrix(enc, &mw, use_w ? W : X, G, N, ld) ||
        !encode_matrix(enc, &mz, use_z ? Z : X, G, N, ld)) {
        fp::set_error("fp_score_signatures: cuTensorMapEncodeTiled failed");
        return FP_ECUDA;
    }
    const StagedPlan p = staged_plan(G, K, nnz);
    char *ws = static_cast<char *>(workspace);
    uint16_t *op = reinterpret_cast<uint16_t *>(ws + p.op_off);
    int32_t *cut = reinterpret_cast<int32_t *>(ws + p.cut_off);
    FP_CHECK_CUDA(cudaMemsetAsync(cut, 0, (size_t)p.rows * (p.n_blocks + 1) * 4, st));
    score_prepare_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(sig_ptr, sig_gene, sig_sign, K, (int)p.n_blocks,
                                                                     op, cut);
    FP_CHECK_LAUNCH("score_prepare_kernel");
    const int n_groups = (int)p.n_groups;
    switch (method) {
        case FP_METHOD_NAIVE:
            return launch_staged<FP_METHOD_NAIVE>(mx, mw, mz, G, N, sig_ptr, op, cut, K, n_groups, gene_mu, out,
                                                  ld_out, st);
        case FP_METHOD_WEIGHTED:
            return launch_staged<FP_METHOD_WEIGHTED>(mx, mw, mz, G, N, sig_ptr, op, cut, K, n_groups, gene_mu, out,
                                                     ld_out, st);
        case FP_METHOD_IMPUTED:
            return launch_staged<FP_METHOD_IMPUTED>(mx, mw, mz, G, N, sig_ptr, op, cut, K, n_groups, gene_mu, out,
                                                    ld_out, st);
        default:
            return launch_staged<FP_METHOD_NONZERO>(mx, mw, mz, G, N, sig_ptr, op, cut, K, n_groups, gene_mu, out,
                                                    ld_out, st);
    }
}
