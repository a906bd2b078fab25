# -*- coding: utf-8 -*-
"""B200 mirror of the scoring half of ``FastProject/Signatures.py``.

Same call surface as the reference (/root/reference/FastProject/Signatures.py):
``Signature``, ``generate_random_sigs`` (:684), ``calculate_sig_scores`` (:630),
``sigs_vs_projections`` (:300), ``p_to_q`` (:536), ``get_bg_dist`` (:19).
All floating-point work on the cells runs on the GPU through the C ABI
(include/fastproject_b200.h); what stays on the host is what must stay
bit-/stream-identical to the reference and is O(signatures): the two global
RNG streams, label bookkeeping, numGenes grouping and the Benjamini-Hochberg
step (whose tie order is NumPy's argsort).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _engine, _native, _scoring
from .Global import RANDOM_SEED
from .SigScoreMethods import SignatureScores

# cache of the permutation background of the precomputed-numeric branch (:16-27)
_bg_dist = np.zeros((0, 0))
# its device twin: (the host array it was made from, [R x n] int64 CUDA tensor).  The host array
# is this module's own state (get_bg_dist replaces it, never edits it), so its identity says
# whether the device copy is current; kept only while it is small (1 GiB).
_bg_dist_dev = None
_BG_DEVICE_CACHE_MAX = 1 << 30


def _drop_bg_device():
    global _bg_dist_dev
    _bg_dist_dev = None


_engine._clear_hooks.append(_drop_bg_device)


def _bg_permutations_device(perm):
    """[R x n] row permutations of ``get_bg_dist`` on the device.  A precomputed signature is
    tested against every projection with the SAME 10 000 permutations: uploading (and
    transposing) 160 MB per call was most of the time of this branch at 2 000 cells."""
    global _bg_dist_dev
    if _bg_dist_dev is not None and _bg_dist_dev[0] is perm:
        return _bg_dist_dev[1]
    dev_perm = _engine.to_device(perm, torch.long).t().contiguous()          # transposed on the device
    _bg_dist_dev = (perm, dev_perm) if perm.nbytes <= _BG_DEVICE_CACHE_MAX else None
    return dev_perm


def get_bg_dist(N_SAMPLES, NUM_REPLICATES):
    """Signatures.py:19-27.  Host RNG on purpose: the draw must equal the
    reference's ``np.random.seed(RANDOM_SEED); rand(N, R)`` stream."""
    global _bg_dist
    if _bg_dist.shape[0] != N_SAMPLES or _bg_dist.shape[1] != NUM_REPLICATES:
        np.random.seed(RANDOM_SEED)
        _bg_dist = np.random.rand(N_SAMPLES, NUM_REPLICATES)
        _bg_dist = np.argsort(_bg_dist, axis=0)
    return _bg_dist


class Signature(object):
    """Signatures.py:730-773."""

    def __init__(self, sig_dict, signed, filename, name):
        self.sig_dict = sig_dict
        self.signed = signed
        self.source = filename
        self.name = name

    def sig_indices(self, genes):
        """(len(genes), 1) float64 vector of {-1, 0, +1} (:738-753).  Kept for
        API compatibility; the scoring path uses the CSR packer instead."""
        out = np.zeros((len(genes), 1), dtype=np.float64)
        rows, signs = _scoring.signature_rows(self.sig_dict, _scoring.gene_row_index(genes))
        out[rows, 0] = signs
        return out


def calculate_sig_scores(data, signatures, method='weighted_avg',
                         zero_locations=None, min_signature_genes=1e99, distributed=False):
    """Signatures.py:630-681.  All signatures (real or random background) are
    scored by ONE kernel launch on the device-resident expression / weight
    matrices; signatures matching no gene or fewer than ``min_signature_genes``
    genes are silently discarded as in the reference (:672-676).

    Returns dict name -> SignatureScores (insertion-ordered).  The score vectors
    stay on the GPU until ``.scores`` / ``.ranks`` is read.

    ``distributed`` (extension): one process per GPU, ONE job -- every rank makes the same call
    with the same data and the same signature list.  The matrices are uploaded cooperatively
    (1/world of the rows per PCIe link, all-gathered over NVLink), a rank matches, packs and
    scores its contiguous block of the signatures, and the returned dictionary names every kept
    signature of the job; reading ``.scores`` / ``.ranks`` of any entry gathers the rows of all
    ranks (a collective: every rank must do it), ``sigs_vs_projections(distributed=True)``
    consumes them on the device."""
    if method not in _scoring.METHODS:
        # the reference leaves sig_score_method unbound here (UnboundLocalError)
        raise UnboundLocalError("unknown signature score method %r" % (method,))
    batch = _scoring.score_batch(data, signatures, method, zero_locations, min_signature_genes,
                                 replicated=distributed)
    sig_scores_dict = dict()
    for row, pos in enumerate(batch.kept):
        sig = signatures[pos]
        sig_scores_dict[sig.name] = SignatureScores._from_batch(
            batch, row, sig.name, data.col_labels, int(batch.num_genes[row]))
    return sig_scores_dict


def prefetch_scores(sig_scores_dict):
    """Extension: start copying the scores of a ``calculate_sig_scores`` result to the host now
    (asynchronously, on a side stream) instead of at the first read of a ``.scores`` -- the copy
    then runs beside ``sigs_vs_projections``.  In a multi-GPU job it covers this rank's rows."""
    seen = set()
    for ss in sig_scores_dict.values():
        batch = getattr(ss, "_batch", None)
        if batch is not None and id(batch) not in seen:
            seen.add(id(batch))
            batch.prefetch_host()


def generate_random_sigs(features, signed, sizes=None, num_per_size=3000):
    """Signatures.py:684-728.  Stays on the host and consumes the same two
    global RNG streams in the same order (``random.sample`` then
    ``np.random.choice`` per signature), so the background is identical to the
    reference's for an identical RNG state."""
    if sizes is None:
        sizes = [5, 10, 20, 50, 100, 200]
    signs = [-1, 1] if signed else [1]
    random_sigs = []
    for size in sizes:
        for j in range(num_per_size):
            new_sig_genes = random.sample(features, size)
            new_sig_signs = np.random.choice(signs, size)
            new_sig_dict = dict()
            for gene, sign in zip(new_sig_genes, new_sig_signs):
                new_sig_dict[gene] = int(sign)
            random_sigs.append(Signature(new_sig_dict, signed, 'x',
                                         name="RANDOM_BG_" + str(size) + "_" + str(j)))
    return random_sigs


def p_to_q(p_values):
    """Benjamini-Hochberg as the reference does it (:536-551): q = p*m/rank with
    rank = argsort(argsort(p)) + 1, no monotone step, clipped to 1.  Host NumPy:
    O(signatures x projections) and the tie order must be argsort's."""
    original_shape = p_values.shape
    p_vals_flat = p_values.flatten()
    rank = p_vals_flat.argsort().argsort() + 1
    q_vals = p_vals_flat * p_values.size / rank
    q_vals.shape = original_shape
    q_vals[q_vals > 1.0] = 1.0
    return q_vals


# --------------------------------------------------------------------------
# sigs_vs_projections
# --------------------------------------------------------------------------
def _device_rank_rows(score_objs, n_cells):
    """Rank rows [len(score_objs) x ld] on the device, in the given order.
    Device-backed SignatureScores are gathered from their batch; host-backed
    ones are uploaded and ranked by the device kernel."""
    dev = _engine.require_cuda()
    ld = _engine.padded(n_cells)
    out = torch.zeros((len(score_objs), ld), dtype=torch.float64, device=dev)
    by_batch = dict()
    host_rows = []
    for pos, ss in enumerate(score_objs):
        batch = getattr(ss, "_batch", None)
        if batch is not None and getattr(ss, "_ranks", None) is None:
            by_batch.setdefault(id(batch), (batch, [], []))
            by_batch[id(batch)][1].append(pos)
            by_batch[id(batch)][2].append(ss._row)
        else:
            host_rows.append(pos)
    for batch, dst, src in by_batch.values():
        ranks = batch.ranks_dev()
        if dst == list(range(dst[0], dst[0] + len(dst))) and src == list(range(src[0], src[0] + len(src))):
            out[dst[0]:dst[0] + len(dst)] = ranks[src[0]:src[0] + len(src)]     # whole batches in order
            continue
        # (index tables go through _engine.to_device: not a copy-engine transfer)
        d = _engine.to_device(np.asarray(dst, dtype=np.int64), torch.long)
        s = _engine.to_device(np.asarray(src, dtype=np.int64), torch.long)
        out.index_copy_(0, d, ranks.index_select(0, s))
    if host_rows:
        need_rank = [p for p in host_rows if getattr(score_objs[p], "_ranks", None) is None]
        if need_rank:
            mat = np.zeros((len(need_rank), ld), dtype=np.float64)
            for r, p in enumerate(need_rank):
                mat[r, :n_cells] = np.asarray(score_objs[p].scores, dtype=np.float64)
            ranked = _engine.rank_rows(_engine.to_device(mat), n_cells)
            out.index_copy_(0, _engine.to_device(np.asarray(need_rank, dtype=np.int64), torch.long), ranked)
            host = ranked[:, :n_cells].cpu().numpy()
            for r, p in enumerate(need_rank):
                score_objs[p]._ranks = host[r]
        have = [p for p in host_rows if p not in set(need_rank)]
        if have:
            mat = np.zeros((len(have), ld), dtype=np.float64)
            for r, p in enumerate(have):
                mat[r, :n_cells] = score_objs[p]._ranks
            out.index_copy_(0, _engine.to_device(np.asarray(have, dtype=np.int64), torch.long),
                            _engine.to_device(mat))
    return out


def _background_groups(random_num_genes, real_num_genes):
    """numGenes bookkeeping of :417-439 (data independent): groups in
    first-seen order; each real signature -> group with the closest numGenes
    (np.argmin: first minimum wins)."""
    group_of = dict()
    members = []
    for k, ng in enumerate(random_num_genes):
        if ng not in group_of:
            group_of[ng] = len(members)
            members.append([])
        members[group_of[ng]].append(k)
    sizes = np.array(list(group_of.keys()), dtype=np.float64)
    order = np.array([k for m in members for k in m], dtype=np.int32)
    ptr = np.zeros(len(members) + 1, dtype=np.int32)
    for g, m in enumerate(members):
        ptr[g + 1] = ptr[g] + len(m)
    if len(real_num_genes) and sizes.size == 0:
        # np.argmin of an empty sequence (:437)
        raise ValueError("attempt to get argmin of an empty sequence")
    sig_group = np.array([int(np.argmin(np.abs(ng - sizes))) for ng in real_num_genes],
                         dtype=np.int32)
    return order, ptr, sig_group


def _factor_design(values):
    """One-hot design of a factor signature (:375-387); level order is
    ``list(set(values))`` as in the reference."""
    levels = list(set(values))
    n = len(values)
    freq = np.zeros(len(levels))
    onehot = np.zeros((len(levels), n))
    for j, lev in enumerate(levels):
        hit = [i for i, v in enumerate(values) if v == lev]
        onehot[j, hit] = 1
        freq[j] = len(hit) / n
    return levels, freq, onehot


def sigs_vs_projections(projections, sig_scores_dict, random_sig_scores_dict,
                        NEIGHBORHOOD_SIZE=0.33, precision=None, sigma_per_cell=None, distributed=False,
                        prefetch_next=None):
    """Evaluates the significance of each signature vs each projection
    (Signatures.py:300-534).

    :param projections: dict of (string) => (numpy.ndarray of shape 2xNum_Samples)
    :param sig_scores_dict: dict of (string) => SignatureScores
    :param random_sig_scores_dict: dict of (string) => SignatureScores (background)
    :param NEIGHBORHOOD_SIZE: Gaussian kernel bandwidth (scalar, as in the reference)
    :param precision: (extension) "auto" | "tc" | "fp64"; default from FASTPROJECT_B200_PRECISION,
        else "auto": the tensor-core contraction (exact digit-sliced int8 products, 32-bit
        fixed-point weights; every value must be a half-integer -- ranks, one-hot columns) and
        the FP64 kernel for anything it refuses or does not cover.  "tc" raises instead of
        falling back; "fp64" always uses the FP64 kernel.
    :param sigma_per_cell: (extension) optional (N,) array of per-cell bandwidths
        replacing NEIGHBORHOOD_SIZE
    :param prefetch_next: (extension) host matrices (expression, weights, ...) of the NEXT data
        set: their upload is started on a side stream once this call has uploaded its own small
        inputs, and overlaps the projection loop (see ``fastproject_b200.prefetch``)
    :param distributed: (extension) one process per GPU (``torch.distributed`` initialised), ONE
        job: every rank passes the same projections and the dictionaries returned by
        ``calculate_sig_scores(..., distributed=True)`` on the same signature lists.  The rank
        rows are all-gathered over NVLink, the (projection x signature-block) grid is cut into
        one run per rank (``distributed.partition_units``), the per-pair medians are summed
        over the ranks (every entry has one owner), and every rank returns the full result --
        what the reference returns for the same call
    :return: (sp_row_labels, sp_col_labels, sig_proj_matrix, sig_proj_matrix_p (log10 q),
              random_sig_proj_matrix, random_sig_score_keys)
    """
    np.random.seed(RANDOM_SEED)                                         # :320
    sp_row_labels, sp_row_labels_factors, sp_row_labels_pnum = [], [], []
    for name, sig_scores in sig_scores_dict.items():                    # :326-333
        if sig_scores.isPrecomputed:
            if sig_scores.isFactor:
                sp_row_labels_factors.append(name)
            else:
                sp_row_labels_pnum.append(name)
        else:
            sp_row_labels.append(name)

    sp_col_labels = list(projections.keys())                            # :337-338
    sp_col_labels.sort()

    N_SAMPLES = len(sig_scores_dict[list(sig_scores_dict.keys())[0]].sample_labels)   # :340-342
    N_SIGNATURES = len(sp_row_labels)
    N_PROJECTIONS = len(sp_col_labels)
    random_sig_score_keys = list(random_sig_scores_dict.keys())         # :366
    N_RANDOM = len(random_sig_score_keys)

    dev = _engine.require_cuda()
    world_size, rank = 1, 0
    if distributed:
        from . import distributed as D
        world_size, rank = D.world()
    std_objs = [sig_scores_dict[k] for k in sp_row_labels]
    rnd_objs = [random_sig_scores_dict[k] for k in random_sig_score_keys]
    ranks = _device_rank_rows(std_objs + rnd_objs, N_SAMPLES)           # :359-369 (all rows, every rank)
    n_rows = N_SIGNATURES + N_RANDOM

    if sigma_per_cell is not None:
        sig2 = _engine.to_device(np.asarray(sigma_per_cell, dtype=np.float64) ** 2)
    else:
        sig2 = NEIGHBORHOOD_SIZE ** 2                                   # :398

    factor_sig_proj_matrix = np.zeros((len(sp_row_labels_factors), N_PROJECTIONS))
    factor_sig_proj_matrix_p = np.zeros((len(sp_row_labels_factors), N_PROJECTIONS))
    pnum_sig_proj_matrix = np.zeros((len(sp_row_labels_pnum), N_PROJECTIONS))
    pnum_sig_proj_matrix_p = np.zeros((len(sp_row_labels_pnum), N_PROJECTIONS))
    factor_dict = dict((s, _factor_design(sig_scores_dict[s].scores)) for s in sp_row_labels_factors)

    for proj in sp_col_labels:
        if tuple(np.shape(projections[proj])) != (2, N_SAMPLES):
            raise ValueError("projection %r has shape %s, expected (2, %d)"
                             % (proj, tuple(np.shape(projections[proj])), N_SAMPLES))
    # all projections in one upload
    coords_all = _engine.to_device(np.stack([np.asarray(projections[p], dtype=np.float64)
                                             for p in sp_col_labels])) if N_PROJECTIONS else None

    if prefetch_next is not None:
        _engine.prefetch(*prefetch_next, replicated=distributed)

    tables = []

    def background_tables():
        # numGenes bookkeeping of :417-439 (host work, O(signatures)): done while the GPU contracts
        order, gptr, sig_group = _background_groups([o.numGenes for o in rnd_objs],
                                                    [o.numGenes for o in std_objs])
        d_tables = _engine.to_device(np.concatenate([order, gptr, sig_group]).astype(np.int32), torch.int32)
        tables.extend([d_tables[:order.size], d_tables[order.size:order.size + gptr.size],
                       d_tables[order.size + gptr.size:]])

    # standard signatures, real and random together (:393-446): this rank's runs of the
    # (projection x signature-block) grid -- on one GPU, every projection with all rows
    if world_size > 1:
        my_runs = D.partition_units(N_PROJECTIONS, n_rows, world_size)[rank]
    else:
        my_runs = [(i, 0, n_rows) for i in range(N_PROJECTIONS)] if n_rows else []
    med_all = torch.zeros((N_PROJECTIONS, n_rows), dtype=torch.float64, device=dev)
    if my_runs:
        overlap = _engine.MedianOverlap(max(hi - lo for _, lo, hi in my_runs), ranks.shape[1], dev)
        _contract_runs(my_runs, coords_all, sig2, ranks, N_SAMPLES, overlap, med_all, precision,
                       while_running=background_tables)
    if not tables:
        background_tables()
    if world_size > 1:
        D.all_reduce_sum(med_all)                                       # the one exchange of stage 2
    p_all = torch.empty((N_PROJECTIONS, N_SIGNATURES), dtype=torch.float64, device=dev)
    if N_SIGNATURES:
        for i in range(N_PROJECTIONS):
            p, _, _ = _engine.background_pvalues(med_all[i, :N_SIGNATURES].contiguous(), tables[2],
                                                 med_all[i, N_SIGNATURES:].contiguous(), tables[0],
                                                 tables[1])            # :417-442
            p_all[i] = p

    for i, proj in enumerate(sp_col_labels):
        coords = coords_all[i]
        for j, sig in enumerate(sp_row_labels_pnum):                    # :451-481
            # multi-GPU: the (signature, projection) pairs of this branch go round the ranks;
            # every rank draws the permutation table (host RNG stream, cached by shape)
            mine = (i * len(sp_row_labels_pnum) + j) % world_size == rank
            cons, pval = _precomputed_numeric(sig_scores_dict[sig], coords, sig2, N_SAMPLES, precision,
                                              compute=mine)
            pnum_sig_proj_matrix[j, i] = cons
            pnum_sig_proj_matrix_p[j, i] = pval

        for j, sig in enumerate(sp_row_labels_factors):                 # :484-518
            # consumes the global NumPy stream: evaluated by every rank in the same order
            cons, pval = _factor(factor_dict[sig], coords, sig2, N_SAMPLES, precision)
            factor_sig_proj_matrix[j, i] = cons
            factor_sig_proj_matrix_p[j, i] = pval

    if world_size > 1 and sp_row_labels_pnum:
        both = _engine.to_device(np.stack([pnum_sig_proj_matrix, pnum_sig_proj_matrix_p]))
        both = D.all_reduce_sum(both).cpu().numpy()
        pnum_sig_proj_matrix, pnum_sig_proj_matrix_p = both[0], both[1]

    med_host = med_all.cpu().numpy()
    sig_proj_matrix = 1 - med_host[:, :N_SIGNATURES].T / N_SAMPLES      # :444
    random_sig_proj_matrix = 1 - med_host[:, N_SIGNATURES:].T / N_SAMPLES   # :445
    sig_proj_matrix_p = p_all.cpu().numpy().T.copy()

    sig_proj_matrix = np.concatenate((sig_proj_matrix, factor_sig_proj_matrix,
                                      pnum_sig_proj_matrix), axis=0)   # :524-526
    sig_proj_matrix_p = np.concatenate((sig_proj_matrix_p, factor_sig_proj_matrix_p,
                                        pnum_sig_proj_matrix_p), axis=0)
    sp_row_labels = sp_row_labels + sp_row_labels_factors + sp_row_labels_pnum

    sig_proj_matrix_p = p_to_q(sig_proj_matrix_p)                       # :528-530
    sig_proj_matrix_p[sig_proj_matrix_p == 0] = 1e-300
    with np.errstate(invalid="ignore", divide="ignore"):
        sig_proj_matrix_p = np.log10(sig_proj_matrix_p)

    return (sp_row_labels, sp_col_labels, sig_proj_matrix, sig_proj_matrix_p,
            random_sig_proj_matrix, random_sig_score_keys)


def _contract_runs(runs, coords_all, sig2, ranks, n, overlap, med_all, precision, while_running=None):
    """med_all[p, lo:hi] = median over cells of |rank - neighbourhood prediction| for every
    run (projection p, rows lo..hi) -- :396-414.  The median of one run is taken on a side stream
    while the next run's weight planes are generated (``_engine.MedianOverlap``).  The tensor-core
    mode refuses values it cannot represent exactly; "auto" then repeats the run on the FP64
    kernel, "tc" raises."""
    requested = precision or _engine.default_precision()
    pending = None                      # (rows view, indices of the runs launched on that shape)

    def settle(batch):
        view, idxs = batch
        if _engine.consistency_refused(view, n, precision=requested, slot="svp"):
            if requested != "auto":
                raise _native.NativeError(
                    "sigs_vs_projections(precision=%r): the values are not exactly representable on the "
                    "tensor-core path (NaN / inf ranks, non-half-integers or too wide a range); use "
                    "precision='fp64' or 'auto'" % (requested,))
            overlap.join()              # the refused runs' (NaN) medians must land before their replacement
            for r in idxs:
                p, lo, hi = runs[r]
                dis = _engine.consistency(coords_all[p], sig2, ranks[lo:hi], n, out=overlap.buffer(hi - lo),
                                          precision="fp64", slot="svp", any_cell_order=True)
                overlap.median_into(dis, n, med_all[p, lo:hi])

    # row ranges contracted against more than one projection (a single GPU: all of them) are
    # packed once -- value range and digit planes -- and only permuted per projection
    uses = dict()
    for _, lo, hi in runs:
        uses[(lo, hi)] = uses.get((lo, hi), 0) + 1
    packed = dict()
    for r, (p, lo, hi) in enumerate(runs):
        view = ranks[lo:hi]
        if pending is not None and pending[0].shape[0] != view.shape[0]:
            settle(pending)             # the flag lives in a workspace laid out for that shape
            pending = None
        if uses[(lo, hi)] > 1 and (lo, hi) not in packed:
            packed[(lo, hi)] = _engine.pack_values(view, n, precision=requested)
        dis = _engine.consistency(coords_all[p], sig2, view, n, out=overlap.buffer(hi - lo),
                                  precision=requested, slot="svp", any_cell_order=True,
                                  packed=packed.get((lo, hi)))                            # only the median is taken
        overlap.median_into(dis, n, med_all[p, lo:hi])
        pending = (view, (pending[1] if pending else []) + [r])
    if while_running is not None:
        while_running()                 # host bookkeeping of the caller, while the queue drains
    if pending is not None:
        settle(pending)
    overlap.join()


def _single_group_pvalue(med_value, bg_medians):
    """p = Phi((m - mean(bg)) / std(bg)) on the device; sigma == 0 -> caller decides."""
    dev = bg_medians.device
    n = bg_medians.shape[0]
    order = torch.arange(n, dtype=torch.int32, device=dev)
    gptr = _engine.to_device(np.array([0, n], dtype=np.int32), torch.int32)
    grp = torch.zeros((1,), dtype=torch.int32, device=dev)
    p, mu, sd = _engine.background_pvalues(med_value.reshape(1), grp, bg_medians, order, gptr)
    return float(p[0]), float(sd[0])


def _precomputed_numeric(sig_scores, coords, sig2, n, precision, compute=True):
    """Precomputed numerical signature vs one projection (:451-481): same
    dissimilarity; background = 10 000 column permutations of the rank vector
    (host RNG stream of get_bg_dist) pushed through the same device kernels.
    ``compute=False`` (another rank owns this pair): only the host RNG side effects."""
    r = np.asarray(sig_scores.ranks, dtype=np.float64)
    if (r == r[0]).all():
        return (0.0, 1.0) if compute else (0.0, 0.0)
    NUM_REPLICATES = 10000
    perm = get_bg_dist(n, NUM_REPLICATES)                # (n, R) row indices
    if not compute:
        return 0.0, 0.0
    dev = coords.device
    ld = _engine.padded(n)
    r_dev = _engine.to_device(r)
    rows = torch.zeros((NUM_REPLICATES + 1, ld), dtype=torch.float64, device=dev)
    rows[0, :n] = r_dev
    rows[1:, :n] = r_dev[_bg_permutations_device(perm)]                      # (R, n)
    dis = _engine.consistency(coords, sig2, rows, n, precision=precision, any_cell_order=True)
    med = _engine.row_medians(dis, n)
    p, sd = _single_group_pvalue(med[0], med[1:])
    if sd == 0:
        p = 1.0
    return 1 - float(med[0]) / n, p


def _factor(design, coords, sig2, n, precision):
    """Factor signature vs one projection (:484-518).  The one-hot and the
    1000 random binary columns go through the device kernel with
    FP_OUT_PREDICTION; the per-cell ``np.random.permutation`` draws stay on the
    host stream and are applied on the device as a gather."""
    levels, freq, onehot = design
    if (freq == 1).any():
        return 0.0, 1.0
    dev = coords.device
    ld = _engine.padded(n)
    n_levels = len(levels)
    NUM_REPLICATES = 1000
    column_assignments = np.random.choice(n_levels, NUM_REPLICATES, p=freq)
    column_assignments = freq[column_assignments].reshape((1, NUM_REPLICATES))
    rand_factors = (np.random.rand(n, NUM_REPLICATES) < column_assignments).astype(np.float64)
    # row r of the reference's random_predictions is shuffled by
    # np.random.permutation(row); draw the equivalent index permutations
    shuffles = np.empty((n, NUM_REPLICATES), dtype=np.int64)
    for ii in range(n):
        shuffles[ii] = np.random.permutation(NUM_REPLICATES)

    rows = torch.zeros((n_levels + NUM_REPLICATES, ld), dtype=torch.float64, device=dev)
    rows[:n_levels, :n] = _engine.to_device(onehot)
    rows[n_levels:, :n] = _engine.to_device(np.ascontiguousarray(rand_factors.T))
    pred = _engine.consistency(coords, sig2, rows, n, output_kind=_native.OUT_PREDICTION,
                               precision=precision)
    oh = rows[:n_levels, :n]
    dissimilarity = torch.zeros((1, ld), dtype=torch.float64, device=dev)
    dissimilarity[0, :n] = 1 - (oh * pred[:n_levels, :n]).sum(dim=0)
    med = _engine.row_medians(dissimilarity, n)

    rp = pred[n_levels:, :n].t().contiguous()                       # (n, R) like the reference
    rp = torch.gather(rp, 1, _engine.to_device(shuffles, torch.long))
    bg = torch.zeros((NUM_REPLICATES, ld), dtype=torch.float64, device=dev)
    bg[:, :n] = (1 - rp).t()
    bg_med = _engine.row_medians(bg, n)
    p, sd = _single_group_pvalue(med[0], bg_med)
    if sd == 0:
        p = 1
    return 1 - float(med[0]), p
