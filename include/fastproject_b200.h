/*
 * fastproject_b200.h -- C ABI of the B200-native FastProject scoring core.
 *
 * The reference (YosefLab/FastProject v1.1.4) is pure Python and has no FFI /
 * plugin layer of its own; the boundary it offers for this path is the Python
 * call surface of FastProject/SigScoreMethods.py and FastProject/Signatures.py.
 * Each entry point below names the reference code it replaces (paths relative
 * to /root/reference/FastProject/).  The Python mirror of that call surface
 * (fastproject_b200/SigScoreMethods.py, Signatures.py) binds these symbols with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host; no torch types, no C++ types, no exceptions;
 *   - every function returns 0 on success or a negative FP_E* code and leaves
 *     a message retrievable with fp_last_error() (thread-local);
 *   - every launch goes to the caller's stream (a cudaStream_t passed as
 *     void*; NULL = legacy default stream); no hidden device allocation:
 *     scratch space is sized by the *_workspace_bytes() query and supplied by
 *     the caller;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - one CUDA device per process (the deployment is one process per GPU): kernel
 *     attributes, the SM count and the resident-cluster count are cached per process.
 *
 * Device data layout
 *   expression X, weights W, zero mask Z : genes x cells, float64, ld >= N
 *       (the reference's ExpressionData.base / .weights / zero_locations,
 *        DataTypes.py:16, Pipelines.py:148)
 *   signatures : CSR over genes -- sig_ptr[K+1] (int32), sig_gene[nnz] (int32,
 *       ascending inside a signature), sig_sign[nnz] (int8, +1/-1)
 *       (replaces the dense G x 1 vector of Signature.sig_indices,
 *        Signatures.py:738-753)
 *   scores / ranks / dissimilarities : signatures x cells, float64 -- one
 *       signature per ROW, cells contiguous (the transpose of the reference's
 *       N x K sig_score_matrix, Signatures.py:359): the layout the segmented
 *       sort, the TMA tiles of the consistency kernel and the per-signature
 *       median all want
 *   coordinates : 2 x N float64 (projections[name], Signatures.py:394)
 */
#ifndef FASTPROJECT_B200_H
#define FASTPROJECT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_ABI_VERSION 3

/* error codes */
#define FP_OK            0
#define FP_EINVAL       -1   /* bad argument                                */
#define FP_ECUDA        -2   /* CUDA runtime / launch error                 */
#define FP_EWORKSPACE   -3   /* workspace too small                         */
#define FP_EUNSUPPORTED -4   /* shape / mode not supported by this build    */

/* scoring methods: Signatures.py:660-667 (method strings of calculate_sig_scores) */
#define FP_METHOD_NAIVE        0   /* SigScoreMethods.py:18-48   "naive"        */
#define FP_METHOD_WEIGHTED     1   /* SigScoreMethods.py:51-82   "weighted_avg" */
#define FP_METHOD_IMPUTED      2   /* SigScoreMethods.py:85-128  "imputed"      */
#define FP_METHOD_NONZERO      3   /* SigScoreMethods.py:131-162 "only_nonzero" */

/* what fp_consistency writes */
#define FP_OUT_ABS_ERROR       0   /* |r - prediction|  (Signatures.py:408, 413, 462, 470) */
#define FP_OUT_PREDICTION      1   /* the prediction itself (factor branch, :493, 505)     */
/* OR-ed into output_kind: the caller only reduces every output row over the cells (the median of
 * Signatures.py:409): the columns of out_dissim may come in the library's own cell order (a
 * Z-order curve on the tensor-core path) instead of the caller's -- same values, same multiset
 * per row.  Without it the columns are in the caller's order (scattered 8-byte writes). */
#define FP_OUT_ANY_CELL_ORDER  0x200

/* arithmetic of the neighbourhood contraction in fp_consistency */
#define FP_PRECISION_FP64      0   /* float64 FMA on the FP64 pipe (verifier)           */
#define FP_PRECISION_TC        1   /* tcgen05 kind::i8: exact digit-sliced contraction   */

int         fp_abi_version(void);
const char *fp_version(void);
const char *fp_last_error(void);

/* Number of kernel launches issued through this library by the calling
 * process so far (all entry points).  bench.py reports the difference over the
 * timed region as "gpu_launches". */
uint64_t    fp_launch_count(void);

/*
 * fp_match_labels -- HOST function (no device work): the row of every queried gene name among
 * the matrix's row labels, or -1.  Both sides are NUL-separated UTF-8 blobs ("\0".join(...)):
 * n_labels labels in labels_bytes bytes, n_queries names in queries_bytes bytes.  Replaces the
 * label-probing loop of Signature.sig_indices (Signatures.py:738-753) for all signatures of a
 * list at once.  Returns FP_OK; 1 if a row label occurs twice (the reference's loop then matches
 * every such row -- the caller takes its per-signature path); FP_EINVAL if a blob does not hold
 * the stated number of names (e.g. a name contains NUL).  out_rows_host: n_queries int32, host.
 */
int fp_match_labels(const char *labels, int64_t labels_bytes, int64_t n_labels,
                    const char *queries, int64_t queries_bytes, int64_t n_queries,
                    int32_t *out_rows_host);

/*
 * fp_gene_detected_mean -- mu_g = sum_c x[g,c]*(1-z[g,c]) / sum_c (1-z[g,c]),
 * NaN -> 0.  Replaces SigScoreMethods.py:109-111 (imputed_eval_signature); the
 * row sums reproduce NumPy's pairwise summation order bit-for-bit.
 * out_mu: G float64.
 */
int fp_gene_detected_mean(const double *X, const double *Z,
                          int64_t G, int64_t N, int64_t ld,
                          double *out_mu, void *stream);

/*
 * fp_score_signatures -- evaluate K signatures on N cells in one launch
 * (real and random-background signatures together).
 * Replaces the per-signature bodies of naive/weighted/imputed/nonzero
 * _eval_signature (SigScoreMethods.py:34-41, 66-75, 101-121, 147-155) and the
 * Python loop of calculate_sig_scores (Signatures.py:669-679).
 * Sums run in ascending gene order with separately rounded multiply and add,
 * so the result is bit-identical to the reference's NumPy float64 evaluation.
 *   W        : required for WEIGHTED and IMPUTED, else may be NULL
 *   Z        : required for IMPUTED and NONZERO, else may be NULL
 *   gene_mu  : required for IMPUTED (from fp_gene_detected_mean), else NULL
 *   out_scores : K x ld_out float64
 *   nnz      : sig_ptr[K] as the host knows it (0 = unknown)
 *   workspace: fp_score_workspace_bytes(G, K, nnz) bytes, or NULL.  nnz and the workspace only
 *              select the data path: with nnz > ceil(K/512) * G, 16-byte aligned matrices and an
 *              even ld the gene rows are staged in shared memory by TMA (cp.async.bulk.tensor),
 *              each (256 genes x 16 cells) box once per group of 512 signatures, and the
 *              workspace holds the membership table re-cut by gene block; otherwise every
 *              signature gathers its rows through L2.  Same arithmetic, same bits.
 * Signatures with zero matched genes must be filtered by the caller (the
 * reference raises ValueError for them, SigScoreMethods.py:60-64).
 */
size_t fp_score_workspace_bytes(int64_t G, int64_t K, int64_t nnz);
int fp_score_signatures(int method,
                        const double *X, const double *W, const double *Z,
                        int64_t G, int64_t N, int64_t ld,
                        const int32_t *sig_ptr, const int32_t *sig_gene,
                        const int8_t *sig_sign, int64_t K, int64_t nnz,
                        const double *gene_mu,
                        double *out_scores, int64_t ld_out,
                        void *workspace, size_t workspace_bytes, void *stream);

/*
 * fp_rank_columns -- average ranks (1-based, ties share the mean rank) of each
 * signature's N scores.  Replaces SignatureScores.ranks =
 * scipy.stats.rankdata(scores, "average") (SigScoreMethods.py:170-178) and
 * the matrix assembly of Signatures.py:359-369.
 *   scores, out_ranks : K x ld float64 (may not alias)
 */
#define FP_RANK_AVERAGE 0   /* scipy rankdata(method="average"): SignatureScores.ranks          */
#define FP_RANK_MIN     1   /* scipy rankdata(method="min"): NormalizationMethods.py:53-61      */
size_t fp_rank_workspace_bytes(int64_t K, int64_t N, int64_t ld);
int fp_rank_columns(const double *scores, int64_t K, int64_t N, int64_t ld,
                    double *out_ranks, void *workspace, size_t workspace_bytes,
                    void *stream);

/* same with the tie rule selectable (FP_RANK_*); rows are ranked independently */
int fp_rank_columns_ex(const double *scores, int64_t K, int64_t N, int64_t ld,
                       double *out_ranks, void *workspace, size_t workspace_bytes,
                       int method, void *stream);

/*
 * fp_compute_weights -- false-negative weights of the expression matrix, the W that
 * fp_score_signatures reads.  Replaces Transforms.compute_weights (Transforms.py:266-316) for
 * the logistic curve fn = L + S / (1 + exp((mu - x0) * a)) that create_false_neg_map (:87-88)
 * fits per cell, and Transforms.estimate_non_detection (:341-381) when p_nd is NULL
 * (p_nd = [x == 0]).
 *   X       : G x N expression (row-major, leading dimension ld)
 *   p_nd    : G x N probability of non-detection (leading dimension ld_pnd) or NULL
 *   params  : 4 x N, rows x0, a, L, S (the reference's params matrix), contiguous
 *   out     : G x N weights in [0, 1] (1 where detected); may not alias X
 * Row sums in NumPy's pairwise order, operations in the reference's order; differs from the
 * reference only by the last-bit freedom of exp() (tested at 1e-12).  NaN propagates as there.
 */
size_t fp_compute_weights_workspace_bytes(int64_t N);
int fp_compute_weights(const double *X, const double *p_nd, const double *params, int64_t G, int64_t N,
                       int64_t ld, int64_t ld_pnd, double *out_weights, int64_t ld_out, void *workspace,
                       size_t workspace_bytes, void *stream);

/*
 * fp_pull_host -- copy `bytes` bytes from PINNED host memory (cudaHostAlloc / cudaHostRegister,
 * addressable by the device under unified addressing) to device memory with a kernel instead
 * of the copy engine.  For the small per-call tables of the path (signature CSR arrays of
 * Signatures.py:669-679, projection coordinates of :391-395): the copy engine serves requests
 * in order, so a 2 MB table issued while the next data set's expression matrices (4.8 GB) are
 * being uploaded would wait for all of it; SM loads over PCIe do not queue behind the engine.
 * dst and src 16-byte aligned.
 */
int fp_pull_host(void *dst_device, const void *src_pinned_host, size_t bytes, void *stream);

/*
 * fp_normalize_rows / fp_normalize_cols -- z-normalisation of every gene (row) or every
 * cell (column) of the genes x cells matrix: (x - mean) / std with population std and
 * std == 0 -> 1.  Replace NormalizationMethods.row_normalization (:34-41, the pipeline
 * default applied at Pipelines.py:202) and col_normalization (:24-31); the sums follow
 * NumPy's order (pairwise along rows, sequential along columns), so the output is
 * bit-identical to the reference.  out may not alias X.  The workspace of the row variant
 * holds the pairwise-summation tree of rows of N elements (built on the device); the
 * calls are asynchronous on the caller's stream.
 */
size_t fp_normalize_rows_workspace_bytes(int64_t N);
int fp_normalize_rows(const double *X, int64_t G, int64_t N, int64_t ld, double *out, int64_t ld_out,
                      void *workspace, size_t workspace_bytes, void *stream);
int fp_normalize_cols(const double *X, int64_t G, int64_t N, int64_t ld, double *out, int64_t ld_out,
                      void *stream);

/*
 * fp_consistency -- for one projection, the local dissimilarity
 *   D[k,i] = | r[k,i] - sum_j w_ij r[k,j] / sum_j w_ij |,
 *   w_ij = exp(-|p_i - p_j|^2 / sigma_sq_i), w_ii = 0, (sum_j w_ij == 0 -> 1)
 * Replaces Signatures.py:396-408 and :412-413 (cdist, exp, zeroed diagonal, row
 * normalisation, both np.dot calls and the |actual - predicted| step) for real and
 * random signatures in one call.  The N x N kernel matrix never exists as doubles:
 * FP_PRECISION_FP64 recomputes it tile by tile; FP_PRECISION_TC keeps it as four 8-bit
 * digit planes (4 bytes per weight) in the caller's workspace, generated once per call
 * in slabs of cells (at most FASTPROJECT_B200_SLAB_BYTES, default 6 GiB, at a time).
 *   coords       : 2 x N float64 (x row then y row)
 *   sigma_sq     : per-cell bandwidth^2 (N float64) or NULL to use sigma_sq_scalar
 *                  (the reference has a single NEIGHBORHOOD_SIZE**2, :398)
 *   ranks        : K x ld float64
 *   out_dissim   : K x ld_out float64
 *   output_kind  : FP_OUT_*  (the factor branch, :493-508, needs the raw prediction)
 *   precision_mode : FP_PRECISION_*
 *   workspace    : fp_consistency_workspace_bytes(N, K, mode) bytes, 1024-byte aligned
 *                  (0 for FP_PRECISION_FP64)
 *
 * FP_PRECISION_TC represents the values as half-integers of bounded range (average ranks,
 * one-hot / binary columns).  Anything else (NaN, non-half-integers, a range wider than
 * three base-256 digits) is refused, not approximated: every output of that call is NaN and a
 * flag is left in the workspace -- fp_consistency_status() reports it to the host.
 */
size_t fp_consistency_workspace_bytes(int64_t N, int64_t K, int precision_mode);

/*
 * fp_consistency_status -- after one or more fp_consistency calls on `workspace` (same
 * N, K, mode): synchronises `stream` and returns FP_OK, or FP_EUNSUPPORTED if the LAST call
 * refused its values (see above; fp_last_error() says why).  The reference has no such
 * state (NumPy float64 takes any value); the Python mirror uses it to raise, or to re-run
 * the call with FP_PRECISION_FP64, instead of handing NaNs to the caller.
 */
int fp_consistency_status(const void *workspace, int64_t N, int64_t K, int precision_mode,
                          void *stream);
int fp_consistency(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                   const double *ranks, int64_t N, int64_t K, int64_t ld,
                   double *out_dissim, int64_t ld_out, int output_kind, int precision_mode,
                   void *workspace, size_t workspace_bytes, void *stream);

/*
 * Packed values -- the values of fp_consistency in the form FP_PRECISION_TC contracts them.
 * sigs_vs_projections tests the SAME rank matrix against every projection
 * (Signatures.py:359-369 builds it once, the loop :388-446 reuses it): fp_values_pack reads the
 * matrix once (value range, balanced base-256 digit planes in the caller's cell order, refusal
 * flag) into a caller-owned buffer, and fp_consistency_packed takes that buffer INSTEAD of the
 * matrix, so a projection only permutes the digit planes into its own cell order.  The buffer
 * is the input: there is no second copy of the values it could disagree with.  Same results,
 * bit for bit, as fp_consistency(FP_PRECISION_TC) on the matrix; fp_consistency_status reports
 * refused values in the same way.
 *   packed : fp_values_pack_bytes(N, K) bytes (0: shape not supported), 256-byte aligned
 */
size_t fp_values_pack_bytes(int64_t N, int64_t K);
int fp_values_pack(const double *values, int64_t K, int64_t N, int64_t ld, void *packed, size_t packed_bytes,
                   void *stream);
int fp_consistency_packed(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                          const void *packed, int64_t N, int64_t K, double *out_dissim, int64_t ld_out,
                          int output_kind, void *workspace, size_t workspace_bytes, void *stream);

/*
 * fp_row_medians -- exact median over the N entries of each of K rows
 * (mean of the two middle order statistics when N is even; NaN if the row
 * holds a NaN).  Replaces np.median(dissimilarity, axis=0), Signatures.py:409,414.
 * Any finite or infinite doubles are accepted (total order on the bit patterns).
 */
int fp_row_medians(const double *values, int64_t K, int64_t N, int64_t ld,
                   double *out_median, void *stream);

/*
 * fp_background_pvalues -- per-numGenes background N(mu, sigma) from the random
 * signatures' medians (population std, ddof = 0, NumPy pairwise summation
 * order) and p = Phi((med - mu) / sigma) for each real signature.
 * Replaces Signatures.py:417-442.  Grouping / nearest-group matching depends
 * only on numGenes and is prepared on the host:
 *   rnd_order[R]         random-signature indices, grouped (group-major, original
 *                        order inside a group)
 *   group_ptr[n_groups+1] extents of each group inside rnd_order
 *   sig_group[K]         for each real signature, the index of its group
 *   out_p[K], out_mu[n_groups], out_sigma[n_groups]
 */
int fp_background_pvalues(const double *med, const int32_t *sig_group, int64_t K,
                          const double *rnd_med, const int32_t *rnd_order,
                          const int32_t *group_ptr, int64_t n_groups,
                          double *out_p, double *out_mu, double *out_sigma,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTPROJECT_B200_H */
