# -*- coding: utf-8 -*-
"""CPU oracle for the FastProject scoring core.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy/SciPy float64, the algorithm of the reference
hot path so that the CUDA product path (``fastproject_b200``) can be checked
against it.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path never falls back
to it.

Parity status: PINNED.  The reference ships no golden vectors for this path
(its test battery asserts no values), so the oracle is
pinned against the live reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the authoring
container, runs it on seeded inputs (and on the reference's own
``Tests/TestFiles`` fixtures) and commits inputs + outputs as ``.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
files bit-for-bit (scores, ranks) or to 1e-12 (BLAS-order dependent values).

Reference lines followed (relative to /root/reference/FastProject/):
  Signature.sig_indices ............ Signatures.py:738-773
  naive_eval_signature ............. SigScoreMethods.py:18-48
  weighted_eval_signature .......... SigScoreMethods.py:51-82
  imputed_eval_signature ........... SigScoreMethods.py:85-128
  nonzero_eval_signature ........... SigScoreMethods.py:131-162
  SignatureScores.ranks ............ SigScoreMethods.py:170-178
  calculate_sig_scores ............. Signatures.py:630-681
  generate_random_sigs ............. Signatures.py:684-728
  sigs_vs_projections .............. Signatures.py:300-534
  get_bg_dist ...................... Signatures.py:19-27
  p_to_q ........................... Signatures.py:536-551
  row/col/row_and_col/col_rank normalisation (the "next" row N1)
  .................................. NormalizationMethods.py:17-61
The arithmetic itself lives in numpy / scipy (rankdata, cdist, norm.cdf),
which are not vendored by the reference (setup.py:19-20 pins only lower
bounds); the golden files record the versions they were generated with.
"""
from __future__ import annotations

import random as _py_random

import numpy as np
from scipy.spatial.distance import cdist
from scipy.stats import norm, rankdata

# Global.py:71 -- "Chosen by roll of a 2147483648-sided die"
RANDOM_SEED = 1335607828

METHOD_NAMES = ("naive", "weighted_avg", "imputed", "only_nonzero")


# --------------------------------------------------------------------------
# containers (attribute contract of DataTypes.ExpressionData, Signature,
# SignatureScores -- only what the hot path touches)
# --------------------------------------------------------------------------
class ExpressionMatrix(object):
    """Stand-in for ``DataTypes.ExpressionData`` (DataTypes.py:16-169): the
    evaluators only read ``.base``, ``.weights``, ``.row_labels``,
    ``.col_labels``."""

    def __init__(self, base, weights, row_labels, col_labels):
        self.base = base
        self.weights = weights
        self.row_labels = list(row_labels)
        self.col_labels = list(col_labels)


class Signature(object):
    """Signatures.py:730-736."""

    def __init__(self, sig_dict, signed, filename, name):
        self.sig_dict = sig_dict
        self.signed = signed
        self.source = filename
        self.name = name

    def sig_indices(self, genes):
        """(G,1) float64 of {-1,0,+1}.  Signatures.py:745-753: a dict value
        equal to 1 or 0 maps to +1, a value equal to -1 maps to -1, anything
        else (e.g. 0.5 from a weighted GMT) leaves 0.  The re-weighting block
        at :761 is dead code (``if(False)``)."""
        column = np.zeros((len(genes), 1), dtype=np.float64)
        members = self.sig_dict
        for pos, gene in enumerate(genes):
            if gene in members:
                v = members[gene]
                if v == 1 or v == 0:
                    column[pos] = 1
                if v == -1:
                    column[pos] = -1
        return column


class SignatureScores(object):
    """SigScoreMethods.py:165-206 (without the JSON writer)."""

    def __init__(self, scores, name, sample_labels, isFactor, isPrecomputed,
                 numGenes):
        self.scores = scores
        self.name = name
        self.sample_labels = sample_labels
        self.isFactor = isFactor
        self.isPrecomputed = isPrecomputed
        self.numGenes = numGenes
        self._ranks = None

    @property
    def ranks(self):
        if self.isFactor:
            raise Exception("Factor signature scores have no rank")
        if self._ranks is None:
            self._ranks = rankdata(self.scores, method="average")
        return self._ranks


# --------------------------------------------------------------------------
# the four evaluators
# --------------------------------------------------------------------------
def _matched_rows(data, signature, min_signature_genes):
    """Common head of all four evaluators (e.g. SigScoreMethods.py:55-64)."""
    column = signature.sig_indices(data.row_labels)
    rows = np.nonzero(column)[0]
    if rows.size == 0:
        raise ValueError("No genes match signature")
    if rows.size < min_signature_genes:
        raise ValueError("Too few genes match signature")
    return rows, column[rows, :]


def _wrap(scores, data, signature, n_matched):
    return SignatureScores(scores, signature.name, data.col_labels,
                           isFactor=False, isPrecomputed=False,
                           numGenes=n_matched)


def naive_eval_signature(data, signature, zeros, min_signature_genes):
    """SigScoreMethods.py:18-48 -- all weights are one."""
    rows, sign = _matched_rows(data, signature, min_signature_genes)
    x = data.base[rows, :]
    ones = np.ones(x.shape)
    scores = (x * sign * ones).sum(axis=0)
    scores /= np.sum(np.abs(sign) * ones, axis=0)
    return _wrap(scores, data, signature, rows.size)


def weighted_eval_signature(data, signature, zeros, min_signature_genes):
    """SigScoreMethods.py:51-82 -- false-negative-weighted average."""
    rows, sign = _matched_rows(data, signature, min_signature_genes)
    w = data.weights[rows, :]
    x = data.base[rows, :]
    scores = (x * sign * w).sum(axis=0)
    den = np.sum(np.abs(sign) * w, axis=0)
    den[den == 0] = 1.0
    scores /= den
    return _wrap(scores, data, signature, rows.size)


def imputed_eval_signature(data, signature, zeros, min_signature_genes):
    """SigScoreMethods.py:85-128 -- zeros replaced by (1-w)*mu_g + w*x."""
    rows, sign = _matched_rows(data, signature, min_signature_genes)
    w = data.weights[rows, :]
    x = data.base[rows, :]
    z = zeros[rows, :]
    nz = 1 - z
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = (x * nz).sum(axis=1, keepdims=True) / nz.sum(axis=1, keepdims=True)
    mu[np.isnan(mu)] = 0.0
    p_expr = 1 - w
    p_not = w
    detected = (x * nz * sign).sum(axis=0)
    filled = np.sum(p_expr * mu * sign * z + p_not * x * sign * z, axis=0)
    scores = detected + filled
    scores /= x.shape[0]
    return _wrap(scores, data, signature, rows.size)


def nonzero_eval_signature(data, signature, zeros, min_signature_genes):
    """SigScoreMethods.py:131-162 -- average over detected entries only."""
    rows, sign = _matched_rows(data, signature, min_signature_genes)
    x = data.base[rows, :]
    z = zeros[rows, :]
    nz = 1 - z
    num = ((x * nz) * sign).sum(axis=0)
    den = nz.sum(axis=0)
    den[den == 0] = 1.0
    return _wrap(num / den, data, signature, rows.size)


_EVALUATORS = {
    "naive": naive_eval_signature,
    "weighted_avg": weighted_eval_signature,
    "imputed": imputed_eval_signature,
    "only_nonzero": nonzero_eval_signature,
}


def calculate_sig_scores(data, signatures, method="weighted_avg",
                         zero_locations=None, min_signature_genes=1e99):
    """Signatures.py:630-681.  A signature whose evaluator raises ValueError
    is silently dropped (:672-676).  An unknown ``method`` leaves the
    evaluator unbound in the reference (UnboundLocalError); here KeyError."""
    evaluator = _EVALUATORS[method]
    out = dict()
    for sig in signatures:
        try:
            out[sig.name] = evaluator(data, sig, zero_locations,
                                      min_signature_genes)
        except ValueError:
            pass
    return out


def generate_random_sigs(features, signed, sizes=None, num_per_size=3000):
    """Signatures.py:684-728.  Consumes the *global* ``random`` stream
    (seeded once at import of Global.py:72) and the global NumPy stream."""
    if sizes is None:
        sizes = [5, 10, 20, 50, 100, 200]
    signs = [-1, 1] if signed else [1]
    out = []
    for size in sizes:
        for j in range(num_per_size):
            genes = _py_random.sample(features, size)
            sgn = np.random.choice(signs, size)
            members = dict()
            for g, s in zip(genes, sgn):
                members[g] = int(s)
            out.append(Signature(members, signed, "x",
                                 name="RANDOM_BG_" + str(size) + "_" + str(j)))
    return out


# --------------------------------------------------------------------------
# signature x projection consistency
# --------------------------------------------------------------------------
_perm_cache = np.zeros((0, 0))


def get_bg_dist(n_samples, n_replicates):
    """Signatures.py:19-27 -- cached column-wise random permutations."""
    global _perm_cache
    if _perm_cache.shape != (n_samples, n_replicates):
        np.random.seed(RANDOM_SEED)
        _perm_cache = np.argsort(np.random.rand(n_samples, n_replicates), axis=0)
    return _perm_cache


def neighbourhood_weights(coords, neighborhood_size, row_slice=None):
    """Row-normalised Gaussian kernel, Signatures.py:396-402.

    ``coords`` is (2,N).  With ``row_slice`` only those rows of the N x N
    matrix are formed (the row-blocked restatement for N where 3*N^2*8 B does
    not fit in RAM); every row is computed by exactly the same operations."""
    pts = coords.T
    rows = pts if row_slice is None else pts[row_slice]
    dist = cdist(rows, pts, metric="euclidean")
    wt = np.exp(-1 * dist ** 2 / neighborhood_size ** 2)
    n = pts.shape[0]
    if row_slice is None:
        wt[np.arange(n), np.arange(n)] = 0
    else:
        r = np.arange(n)[row_slice]
        wt[np.arange(r.size), r] = 0
    tot = np.sum(wt, axis=1, keepdims=True)
    tot[tot == 0] = 1.0
    wt /= tot
    return wt


def _median_abs_error(wt, values):
    """median_i |v_ik - (wt @ v)_ik|, Signatures.py:404-409."""
    pred = np.dot(wt, values)
    return np.median(np.abs(values - pred), axis=0)


def background_table(random_medians, random_num_genes):
    """Per-numGenes (mu, sigma) table in first-seen order, :417-430."""
    groups = dict()
    for v, ng in zip(random_medians, random_num_genes):
        groups.setdefault(ng, []).append(v)
    table = np.zeros((len(groups), 3))
    for row, ng in enumerate(groups.keys()):
        table[row, 0] = ng
        table[row, 1] = np.mean(groups[ng])
        table[row, 2] = np.std(groups[ng])
    return table


def background_pvalues(medians, num_genes, table):
    """Nearest-numGenes lookup (argmin, first wins) + normal CDF, :432-442."""
    mu = np.zeros(medians.shape)
    sd = np.zeros(medians.shape)
    for k in range(medians.size):
        row = np.argmin(np.abs(num_genes[k] - table[:, 0]))
        mu[k] = table[row, 1]
        sd[k] = table[row, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        return norm.cdf((medians - mu) / sd)


def p_to_q(p_values):
    """Benjamini-Hochberg without the monotone step, Signatures.py:536-551."""
    flat = p_values.flatten()
    order = flat.argsort().argsort() + 1
    q = flat * p_values.size / order
    q.shape = p_values.shape
    q[q > 1.0] = 1.0
    return q


def _factor_design(values):
    """One-hot matrix + level frequencies, Signatures.py:375-387.  Level
    order is ``list(set(values))`` -- hash order, as in the reference."""
    levels = list(set(values))
    n = len(values)
    freq = np.zeros(len(levels))
    onehot = np.zeros((n, len(levels)))
    for j, lev in enumerate(levels):
        hit = [i for i, v in enumerate(values) if v == lev]
        onehot[hit, j] = 1
        freq[j] = len(hit) / n
    return levels, freq, onehot


def sigs_vs_projections(projections, sig_scores_dict, random_sig_scores_dict,
                        NEIGHBORHOOD_SIZE=0.33, row_block=None):
    """Signatures.py:300-534.  ``row_block`` (oracle-only) evaluates the
    kernel matrix in row blocks of that many cells; the per-row arithmetic is
    unchanged, only BLAS summation order may differ (<=1e-12)."""
    np.random.seed(RANDOM_SEED)                                    # :320
    std_names, factor_names, pnum_names = [], [], []
    for name, ss in sig_scores_dict.items():                       # :326-333
        if ss.isPrecomputed:
            (factor_names if ss.isFactor else pnum_names).append(name)
        else:
            std_names.append(name)
    proj_names = sorted(projections.keys())                        # :337-338
    first = sig_scores_dict[list(sig_scores_dict.keys())[0]]
    n = len(first.sample_labels)                                   # :340-342
    n_proj = len(proj_names)

    ranks = np.zeros((n, len(std_names)))
    for j, name in enumerate(std_names):
        ranks[:, j] = sig_scores_dict[name].ranks                  # :361-362
    rnd_names = list(random_sig_scores_dict.keys())
    rnd_ranks = np.zeros((n, len(rnd_names)))
    for j, name in enumerate(rnd_names):
        rnd_ranks[:, j] = random_sig_scores_dict[name].ranks       # :368-369
    rnd_num_genes = [random_sig_scores_dict[k].numGenes for k in rnd_names]
    std_num_genes = [sig_scores_dict[k].numGenes for k in std_names]

    cons = np.zeros((len(std_names), n_proj))
    pval = np.zeros((len(std_names), n_proj))
    f_cons = np.zeros((len(factor_names), n_proj))
    f_pval = np.zeros((len(factor_names), n_proj))
    q_cons = np.zeros((len(pnum_names), n_proj))
    q_pval = np.zeros((len(pnum_names), n_proj))
    rnd_cons = np.zeros((len(rnd_names), n_proj))

    designs = dict((s, _factor_design(sig_scores_dict[s].scores))
                   for s in factor_names)

    def row_blocks():
        if row_block is None:
            yield slice(None)
        else:
            for lo in range(0, n, row_block):
                yield slice(lo, min(n, lo + row_block))

    for i, proj in enumerate(proj_names):
        coords = projections[proj]
        if row_block is None:
            wt = neighbourhood_weights(coords, NEIGHBORHOOD_SIZE)
            med = _median_abs_error(wt, ranks)
            rmed = _median_abs_error(wt, rnd_ranks)
        else:
            if factor_names or pnum_names:
                raise NotImplementedError("row-blocked oracle covers the "
                                          "standard-signature branch only")
            dis = np.empty_like(ranks)
            rdis = np.empty_like(rnd_ranks)
            for sl in row_blocks():
                wt = neighbourhood_weights(coords, NEIGHBORHOOD_SIZE, sl)
                dis[sl] = np.abs(ranks[sl] - np.dot(wt, ranks))
                rdis[sl] = np.abs(rnd_ranks[sl] - np.dot(wt, rnd_ranks))
            med = np.median(dis, axis=0)
            rmed = np.median(rdis, axis=0)

        table = background_table(rmed, rnd_num_genes)              # :417-430
        pval[:, i] = background_pvalues(med, std_num_genes, table)  # :432-442
        cons[:, i] = 1 - med / n                                   # :444
        rnd_cons[:, i] = 1 - rmed / n                              # :445

        for j, name in enumerate(pnum_names):                      # :451-481
            r = sig_scores_dict[name].ranks
            if (r == r[0]).all():
                q_cons[j, i] = 0.0
                q_pval[j, i] = 1.0
                continue
            r = r.reshape((r.size, 1))
            m = np.median(np.abs(r - np.dot(wt, r)))
            perm = get_bg_dist(n, 10000)
            bg = r.ravel()[perm]
            bg_med = np.median(np.abs(bg - np.dot(wt, bg)), axis=0)
            mu, sd = np.mean(bg_med), np.std(bg_med)
            q_pval[j, i] = norm.cdf((m - mu) / sd) if sd != 0 else 1.0
            q_cons[j, i] = 1 - m / n

        for j, name in enumerate(factor_names):                    # :484-518
            levels, freq, onehot = designs[name]
            if (freq == 1).any():
                f_cons[j, i] = 0.0
                f_pval[j, i] = 1.0
                continue
            pred = np.dot(wt, onehot)
            m = np.median(1 - np.sum(onehot * pred, axis=1))
            pick = np.random.choice(len(levels), 1000, p=freq)
            thresh = freq[pick].reshape((1, 1000))
            fake = (np.random.rand(n, 1000) < thresh).astype("int")
            fake_pred = np.dot(wt, fake)
            for r_i in range(fake_pred.shape[0]):
                fake_pred[r_i] = np.random.permutation(fake_pred[r_i])
            bg_med = np.median(1 - fake_pred, axis=0)
            mu, sd = np.mean(bg_med), np.std(bg_med)
            f_pval[j, i] = 1 if sd == 0 else norm.cdf((m - mu) / sd)
            f_cons[j, i] = 1 - m

    cons = np.concatenate((cons, f_cons, q_cons), axis=0)          # :524-526
    pval = np.concatenate((pval, f_pval, q_pval), axis=0)
    row_names = std_names + factor_names + pnum_names
    qval = p_to_q(pval)                                            # :528-530
    qval[qval == 0] = 1e-300
    with np.errstate(invalid="ignore", divide="ignore"):
        qval = np.log10(qval)
    return (row_names, proj_names, cons, qval, rnd_cons, rnd_names)


# --------------------------------------------------------------------------
# pieces of the surrounding pipeline needed to build realistic inputs
# --------------------------------------------------------------------------
def row_normalization(data):
    """NormalizationMethods.py:34-41 -- per-gene z-score (sigma 0 -> 1)."""
    mu = data.mean(axis=1, keepdims=True)
    sd = data.std(axis=1, keepdims=True)
    sd[sd == 0] = 1.0
    return (data - mu) / sd


def col_normalization(data):
    """NormalizationMethods.py:24-31 -- per-cell z-score (sigma 0 -> 1)."""
    mu = data.mean(axis=0, keepdims=True)
    sd = data.std(axis=0, keepdims=True)
    sd[sd == 0] = 1.0
    return (data - mu) / sd


def row_and_col_normalization(data):
    """NormalizationMethods.py:44-50 -- rows, then columns."""
    return col_normalization(row_normalization(data))


def col_rank_normalization(data):
    """NormalizationMethods.py:53-61 -- column-wise ranks, ties take the smallest rank."""
    out = np.zeros(data.shape)
    for i in range(data.shape[1]):
        out[:, i] = rankdata(data[:, i], method="min")
    return out


def logistic_fnr(xvals, x0, a, L=0, S=1):
    """Transforms.py:87-88 -- the false-negative-rate curve that create_false_neg_map fits per
    cell and hands to compute_weights as ``fit_func``."""
    return L + S / (1 + np.exp((xvals - x0) * a))


def estimate_non_detection(data):
    """Transforms.py:341-381 with DO_GMM = False (:358): probability of non-detection is the
    zero mask."""
    return (np.asarray(data) == 0).astype("float")


def compute_weights(fit_func, params, data, p_nd):
    """Transforms.py:266-316 -- weights p(not expressed | not detected) of the zero entries,
    1.0 for detected values.  ``data``, ``p_nd``: genes x cells arrays; ``params``: 4 x cells."""
    data = np.asarray(data, dtype=np.float64)
    p_nd = np.array(p_nd, dtype=np.float64)
    p_d = 1 - p_nd
    with np.errstate(all="ignore"):
        mu_h = (data * p_d).sum(axis=1) / p_d.sum(axis=1)
        fn_prob = np.zeros(data.shape)
        for i in range(fn_prob.shape[1]):
            fn_prob[:, i] = fit_func(mu_h, *params[:, i]).ravel()
        pd_e = 1 - fn_prob
        pnd = p_nd.mean(axis=1, keepdims=True)
        pe = (1 - pnd) / (pd_e).mean(axis=1, keepdims=True)
        # (:305 of the reference compares instead of assigning: NaN stays NaN)
        pnd[pnd == 0] = 1.0 / data.shape[1]
        pne_nd = 1 - (1 - pd_e) * pe / pnd
        pne_nd[pne_nd < 0] = 0.0
        pne_nd[pne_nd > 1] = 1.0
        return pne_nd * p_nd + p_d


def normalize_projection(xy):
    """Projections.py:100-113 -- centre, scale 90th-percentile radius to 1.
    ``xy`` is (N,2); returns (2,N)."""
    xy = np.array(xy, dtype=np.float64)
    xy[:, 0] = xy[:, 0] - np.mean(xy[:, 0])
    xy[:, 1] = xy[:, 1] - np.mean(xy[:, 1])
    r = np.sum(xy ** 2, axis=1) ** 0.5
    r90 = np.percentile(r, 90)
    if r90 > 0:
        xy /= r90
    return xy.T
