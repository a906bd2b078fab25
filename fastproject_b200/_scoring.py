# -*- coding: utf-8 -*-
"""Host-side logic of the scoring half of the path: signature -> CSR packing
(replacing the O(G) Python loop of Signature.sig_indices,
/root/reference/FastProject/Signatures.py:738-753), the device-resident
``ScoreBatch`` and its lazy host views."""
from __future__ import annotations

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


def pack_signatures(signatures, row_labels, min_signature_genes, raise_on_reject=False):
    """list[Signature] -> (kept positions, sig_ptr, sig_gene, sig_sign, num_genes).
    Signatures with no matched gene, or fewer than ``min_signature_genes``, are
    rejected -- the reference raises ValueError for them
    (SigScoreMethods.py:60-64) and calculate_sig_scores drops them (:672-676).

    All signatures are packed in one call of the host extension (csrc/pack_ext.c): it walks the
    sig_dicts, looks every gene up in a table of the row labels (the strings' cached hashes) and
    writes the CSR arrays directly.  Duplicated or non-string row labels, non-numeric dict values
    and ``raise_on_reject`` take the per-signature path."""
    table = None if raise_on_reject else _label_table(row_labels)
    packed = None
    if table is not None:
        from . import _native
        packed = _native.pack_ext().pack([sig.sig_dict for sig in signatures], table, float(min_signature_genes))
    if packed is None:
        return _pack_signatures_slow(signatures, gene_row_index(row_labels), min_signature_genes, raise_on_reject)
    kept, ptr, rows, signs, counts = packed
    return (np.frombuffer(kept, dtype=np.int64).tolist(), np.frombuffer(ptr, dtype=np.int32),
            np.frombuffer(rows, dtype=np.int32), np.frombuffer(signs, dtype=np.int8),
            np.frombuffer(counts, dtype=np.int64))


_table_cache = []        # [(copy of the label list, table or None)], most recent first


def _label_table(row_labels):
    """The extension's label table of ``row_labels``; the last few are kept, and one is reused only
    for a list that compares equal to the copy it was built from (a list comparison by identity
    of the elements: microseconds) -- calculate_sig_scores is called twice per data set (real and
    background signatures) with the same labels."""
    if not isinstance(row_labels, list):
        row_labels = list(row_labels)
    for pos, (labels, table) in enumerate(_table_cache):
        if len(labels) == len(row_labels) and labels == row_labels:
            if pos:
                _table_cache.insert(0, _table_cache.pop(pos))
            return table
    from . import _native
    table = _native.pack_ext().label_table(row_labels)
    _table_cache.insert(0, (list(row_labels), table))
    del _table_cache[3:]
    return table


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
        self._host_local_pending = None

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

    def prefetch_host(self):
        """Start the device -> host copy of this rank's score rows now, on a side stream: it runs
        beside the kernels queued after it (``Signatures.prefetch_scores``) and
        ``host_scores_local`` / ``host_scores`` then only wait for it."""
        if self._host_local_pending is None and self._host_scores is None and self.scores_dev.shape[0]:
            self._host_local_pending = _engine.to_host_async(self.scores_dev[:, :self.n_cells])

    def host_scores_local(self):
        """This rank's rows only (no collective), in the order of ``local_rows``."""
        if self._host_local_pending is not None:
            host, done = self._host_local_pending
            done.synchronize()
            return host.numpy()
        if self.owner_rows is None and self._host_scores is not None:
            return self._host_scores
        return _engine.to_host(self.scores_dev[:, :self.n_cells])

    def host_scores(self):
        if self._host_scores is None:
            if self.owner_rows is None:
                self._host_scores = self.host_scores_local()
            else:
                from . import distributed as D
                dev = D.all_gather_owned_rows(self.scores_dev, self.owner_rows)
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
