# -*- coding: utf-8 -*-
"""Host-side logic of the scoring half of the path: signature -> CSR packing
(replacing the O(G) Python loop of Signature.sig_indices,
/root/reference/FastProject/Signatures.py:738-753), the device-resident
``ScoreBatch`` and its lazy host views."""
from __future__ import annotations

import itertools
import operator

import numpy as np
import torch

from . import _engine

METHODS = ("naive", "weighted_avg", "imputed", "only_nonzero")


def gene_row_index(row_labels):
    """gene label -> row (int) or rows (list) if the label is duplicated.  The
    reference loop (:747-753) visits every row, so duplicates all match."""
    index = dict(zip(row_labels, range(len(row_labels))))
    if len(index) == len(row_labels):        # no duplicated label: the common case, one C-level pass
        return index
    index = dict()
    for pos, gene in enumerate(row_labels):
        prev = index.get(gene)
        if prev is None:
            index[gene] = pos
        elif isinstance(prev, list):
            prev.append(pos)
        else:
            index[gene] = [prev, pos]
    return index


def signature_rows(sig_dict, index):
    """Matched (rows ascending, signs) of one signature.  A dict value equal to
    1 or 0 counts as +1, equal to -1 as -1, anything else is ignored (:750-753)."""
    rows, signs = [], []
    for gene, val in sig_dict.items():
        hit = index.get(gene)
        if hit is None:
            continue
        if val == 1 or val == 0:
            s = 1
        elif val == -1:
            s = -1
        else:
            continue
        if isinstance(hit, list):
            rows.extend(hit)
            signs.extend([s] * len(hit))
        else:
            rows.append(hit)
            signs.append(s)
    if len(rows) > 1:
        order = np.argsort(np.asarray(rows, dtype=np.int64), kind="stable")
        rows = np.asarray(rows, dtype=np.int32)[order]
        signs = np.asarray(signs, dtype=np.int8)[order]
    else:
        rows = np.asarray(rows, dtype=np.int32)
        signs = np.asarray(signs, dtype=np.int8)
    return rows, signs


def _match_rows(flat_genes, n_entries, row_labels):
    """Rows of the named genes (-1: not on the matrix) through the library's host-side hash
    table (``fp_match_labels``); None when the labels are duplicated or are not plain strings
    (the caller then takes the per-signature path)."""
    from . import _native
    try:
        labels = "\0".join(row_labels).encode("utf-8")
        queries = "\0".join(flat_genes).encode("utf-8")
    except TypeError:
        return None
    rows = np.empty(n_entries, dtype=np.int32)
    rc = _native.load().fp_match_labels(labels, len(labels), len(row_labels), queries, len(queries),
                                        n_entries, rows.ctypes.data)
    if rc != 0:
        return None
    return rows


def _signs_of(dicts, n_entries):
    """+1 for a dict value equal to 1 or 0, -1 for -1, 0 (= ignored) for anything else
    (Signatures.py:750-753); None if the values are not numbers."""
    kinds = set(itertools.chain.from_iterable(map(dict.values, dicts)))
    try:
        if kinds <= {1, 0}:                       # unsigned lists, every random background list
            return np.ones(n_entries, dtype=np.int8)
        flat_vals = list(itertools.chain.from_iterable(map(dict.values, dicts)))
        if kinds <= {1, 0, -1}:
            v = np.frombuffer(bytes(map((1).__add__, flat_vals)), dtype=np.uint8).astype(np.int8) - 1
            return np.where(v == 0, 1, v).astype(np.int8)
        vals = np.fromiter(flat_vals, dtype=np.float64, count=len(flat_vals))
    except (TypeError, ValueError):
        return None
    return np.where((vals == 1) | (vals == 0), 1, np.where(vals == -1, -1, 0)).astype(np.int8)


def pack_signatures(signatures, row_labels, min_signature_genes, raise_on_reject=False):
    """list[Signature] -> (kept positions, sig_ptr, sig_gene, sig_sign, num_genes).
    Signatures with no matched gene, or fewer than ``min_signature_genes``, are
    rejected -- the reference raises ValueError for them
    (SigScoreMethods.py:60-64) and calculate_sig_scores drops them (:672-676).

    All signatures are matched in one pass: the gene names of the whole list go to the
    library's host-side hash table as one NUL-separated blob (no Python object per gene), one
    sort orders the entries by (signature, gene row); duplicated row labels and non-string
    names take the per-signature path."""
    if raise_on_reject or len(signatures) < 8:
        return _pack_signatures_slow(signatures, gene_row_index(row_labels), min_signature_genes, raise_on_reject)
    dicts = [sig.sig_dict for sig in signatures]
    lengths = np.fromiter((len(d) for d in dicts), dtype=np.int64, count=len(dicts))
    n_entries = int(lengths.sum())
    rows = _match_rows(itertools.chain.from_iterable(dicts), n_entries, row_labels) if n_entries else \
        np.zeros(0, dtype=np.int32)
    signs = _signs_of(dicts, n_entries) if rows is not None else None
    if rows is None or signs is None:
        return _pack_signatures_slow(signatures, gene_row_index(row_labels), min_signature_genes, raise_on_reject)
    rows = rows.astype(np.int64)
    sig_id = np.repeat(np.arange(len(dicts), dtype=np.int64), lengths)
    ok = (rows >= 0) & (signs != 0)
    if not ok.all():
        rows, signs, sig_id = rows[ok], signs[ok], sig_id[ok]
    counts_all = np.bincount(sig_id, minlength=len(dicts)).astype(np.int64)
    keep_sig = (counts_all > 0) & (counts_all >= min_signature_genes)
    if not keep_sig.all():
        entry_ok = keep_sig[sig_id]
        rows, signs, sig_id = rows[entry_ok], signs[entry_ok], sig_id[entry_ok]
    # by signature, ascending gene row inside: one sort of a combined key carrying the sign bit
    # (32-bit keys sort about twice as fast; they hold 2 * signatures * genes < 2^31)
    n_rows_total = len(row_labels)
    key_t = np.int32 if 2 * len(dicts) * n_rows_total < 2 ** 31 else np.int64
    key = ((sig_id * n_rows_total + rows) * 2 + (signs > 0)).astype(key_t)
    key.sort()
    signs = ((key & 1) * 2 - 1).astype(np.int8)
    rows = (key >> 1) % n_rows_total
    kept = np.nonzero(keep_sig)[0]
    counts = counts_all[kept]
    ptr = np.zeros(len(kept) + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    if ptr[-1] >= 2 ** 31:
        raise ValueError("signature membership table exceeds int32 indexing")
    return ([int(k) for k in kept], ptr.astype(np.int32), rows.astype(np.int32),
            signs.astype(np.int8), counts)


def _pack_signatures_slow(signatures, index, min_signature_genes, raise_on_reject=False):
    kept, ptr, genes, signs, counts = [], [0], [], [], []
    for pos, sig in enumerate(signatures):
        rows, sg = signature_rows(sig.sig_dict, index)
        if rows.size == 0:
            if raise_on_reject:
                raise ValueError("No genes match signature")
            continue
        if rows.size < min_signature_genes:
            if raise_on_reject:
                raise ValueError("Too few genes match signature")
            continue
        kept.append(pos)
        genes.append(rows)
        signs.append(sg)
        counts.append(rows.size)
        ptr.append(ptr[-1] + rows.size)
    if ptr[-1] >= 2 ** 31:
        raise ValueError("signature membership table exceeds int32 indexing")
    sig_ptr = np.asarray(ptr, dtype=np.int32)
    sig_gene = np.concatenate(genes) if genes else np.zeros((0,), dtype=np.int32)
    sig_sign = np.concatenate(signs) if signs else np.zeros((0,), dtype=np.int8)
    return kept, sig_ptr, sig_gene, sig_sign, np.asarray(counts, dtype=np.int64)


class ScoreBatch(object):
    """Scores of many signatures on the same cells, resident on the device as a
    [signatures x ld] float64 matrix (one signature per row).

    In a multi-GPU job (``owner_rows`` given) this rank holds only the rows
    ``owner_rows[rank]`` of the job's batch; ``ranks_dev`` / ``host_scores`` / ``host_ranks``
    return ALL rows in job order and contain one all-gather over NVLink the first time they are
    used -- every rank must reach them in the same order (``sigs_vs_projections`` does)."""

    def __init__(self, scores_dev, n_cells, num_genes, owner_rows=None, rank=0):
        self.scores_dev = scores_dev          # local rows
        self.n_cells = int(n_cells)
        self.num_genes = num_genes            # all rows of the job
        self.owner_rows = owner_rows
        self.rank = rank
        self._ranks_dev = None
        self._host_scores = None
        self._host_ranks = None

    @property
    def n_rows(self):
        if self.owner_rows is not None:
            return sum(len(r) for r in self.owner_rows)
        return self.scores_dev.shape[0]

    @property
    def local_rows(self):
        """Job rows held by this rank (ascending)."""
        if self.owner_rows is not None:
            return self.owner_rows[self.rank]
        return np.arange(self.scores_dev.shape[0], dtype=np.int64)

    def local_ranks_dev(self):
        return _engine.rank_rows(self.scores_dev, self.n_cells)

    def ranks_dev(self):
        if self._ranks_dev is None:
            local = self.local_ranks_dev()
            if self.owner_rows is not None:
                from . import distributed as D
                local = D.all_gather_rank_rows(local, self.owner_rows, self.n_cells)
            self._ranks_dev = local
        return self._ranks_dev

    def host_scores_local(self):
        """This rank's rows only (no collective), in the order of ``local_rows``."""
        return _engine.to_host(self.scores_dev[:, :self.n_cells])

    def host_scores(self):
        if self._host_scores is None:
            dev = self.scores_dev
            if self.owner_rows is not None:
                from . import distributed as D
                dev = D.all_gather_owned_rows(dev, self.owner_rows)
            self._host_scores = _engine.to_host(dev[:, :self.n_cells])
        return self._host_scores

    def host_ranks(self):
        if self._host_ranks is None:
            self._host_ranks = _engine.to_host(self.ranks_dev()[:, :self.n_cells])
        return self._host_ranks


def _matrix_of(data):
    base = getattr(data, "base", None)
    if base is None or (isinstance(data, np.ndarray) and not hasattr(data, "row_labels")):
        base = data
    return base


def score_batch(data, signatures, method, zero_locations, min_signature_genes,
                raise_on_reject=False, replicated=False):
    """Score ``signatures`` on ``data``; returns a ScoreBatch whose ``kept``
    lists the positions (into ``signatures``) that survived.

    ``replicated`` (multi-GPU job: every rank makes this call with the same arguments): the
    signature list is dealt round-robin by position to the ranks; a rank matches, packs and
    scores its share only, and the kept positions / numGenes of all shares are exchanged (one
    small all-reduce over the host group), so every rank knows the job's rows."""
    if method not in METHODS:
        raise KeyError(method)
    _engine.require_cuda()
    # the (asynchronous, pinned-memory) uploads are enqueued first so that the host-side CSR
    # packing below runs while the matrices cross PCIe
    x = _engine.expression_cache.get(_matrix_of(data), replicated)
    n_genes, n_cells = x.shape
    w = z = None
    if method in ("weighted_avg", "imputed"):
        w = _engine.expression_cache.get(data.weights, replicated)
        if tuple(w.shape) != (n_genes, n_cells):
            raise ValueError("weights shape %s != data shape %s" % (tuple(w.shape), (n_genes, n_cells)))
    if method in ("imputed", "only_nonzero"):
        if zero_locations is None:
            raise TypeError("method %r needs zero_locations" % method)
        z = _engine.expression_cache.get(zero_locations, replicated)
        if tuple(z.shape) != (n_genes, n_cells):
            # the kernels index Z with X's rows and leading dimension; the reference fails on the
            # fancy-indexing / broadcast of a mismatched mask (SigScoreMethods.py:101-103, 147)
            raise ValueError("zero_locations shape %s != data shape %s" % (tuple(z.shape), (n_genes, n_cells)))
    owner_rows, rank, world_size = None, 0, 1
    mine = signatures
    if replicated:
        from . import distributed as D
        world_size, rank = D.world()
        if world_size > 1:
            mine = signatures[rank::world_size]
    kept, sig_ptr, sig_gene, sig_sign, counts = pack_signatures(
        mine, data.row_labels, min_signature_genes, raise_on_reject)
    # one upload for the two int32 tables (few, larger host -> device transfers: they share the
    # copy engine with a prefetch of the next data set)
    d_tables = _engine.to_device(np.concatenate([sig_ptr, sig_gene]), torch.int32)
    d_ptr, d_gene = d_tables[:sig_ptr.size], d_tables[sig_ptr.size:]
    d_sign = _engine.to_device(sig_sign, torch.int8)
    scores = _engine.score_signatures(method, x, w, z, d_ptr, d_gene, d_sign, len(kept))
    if world_size > 1:
        # numGenes by position (0 = rejected), summed over the ranks: the job's kept rows
        from . import distributed as D
        by_pos = np.zeros(len(signatures), dtype=np.int64)
        by_pos[np.asarray(kept, dtype=np.int64) * world_size + rank] = counts
        # through the host (gloo) group: the launch queue of the GPU keeps running meanwhile
        by_pos = D.all_reduce_sum_host(by_pos)
        all_kept = np.nonzero(by_pos)[0]
        owner_rows = D.cyclic_owner_rows(all_kept.size, world_size, positions=all_kept)
        kept, counts = [int(k) for k in all_kept], by_pos[all_kept]
    batch = ScoreBatch(scores, n_cells, counts, owner_rows, rank)
    batch.kept = kept
    return batch


def rank_vector(values):
    """Average ranks of one host vector through the device kernel."""
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
    n = v.size
    ld = _engine.padded(n)
    dev = torch.zeros((1, ld), dtype=torch.float64, device=_engine.require_cuda())
    dev[0, :n] = _engine.to_device(v)
    return _engine.rank_rows(dev, n)[0, :n].cpu().numpy()
