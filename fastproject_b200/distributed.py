# -*- coding: utf-8 -*-
"""Multi-GPU partitioning of the scoring core: one process per GPU, the
signature rows (real + random background) are split into contiguous blocks,
one block per rank.  Scoring, ranking, the neighbourhood contraction and the
medians of a block need nothing from the other blocks, so the data path has no
collective.  The only exchange is the all-gather of the per-(projection,
signature) median dissimilarities -- O(signatures x projections) doubles --
which every rank needs because the background N(mu, sigma) of a numGenes group
(/root/reference/FastProject/Signatures.py:417-430) pools the random
signatures of ALL blocks.

The reference has no counterpart (it is single-process); the functions below
only decide who owns which rows and put the gathered medians back into the
reference's row order.  They work on any ``torch.distributed`` backend (NCCL
on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_bounds(n_rows, world_size, rank):
    """Contiguous, balanced block [lo, hi) of ``n_rows`` rows owned by ``rank``
    (the first ``n_rows % world_size`` ranks get one extra row)."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(local, n_rows_total=None, group=None):
    """local: [P x rows_local] tensor of this rank's block ->
    [P x rows_total] with the blocks in rank order.  Blocks may have unequal
    widths (they are padded to the widest for the collective)."""
    world_size, rank = world()
    if world_size == 1:
        return local
    widths = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    all_w = [torch.zeros_like(widths) for _ in range(world_size)]
    dist.all_gather(all_w, widths, group=group)
    all_w = [int(w[0]) for w in all_w]
    wmax = max(all_w)
    padded = local
    if local.shape[1] != wmax:
        padded = torch.zeros((local.shape[0], wmax), dtype=local.dtype, device=local.device)
        padded[:, :local.shape[1]] = local
    padded = padded.contiguous()
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    out = torch.cat([p[:, :w] for p, w in zip(parts, all_w)], dim=1)
    if n_rows_total is not None and out.shape[1] != n_rows_total:
        raise RuntimeError("gathered %d rows, expected %d" % (out.shape[1], n_rows_total))
    return out


def _small_to(arr, device):
    """Small host table -> tensor on ``device``; on CUDA through the engine's kernel fetch, so
    that it does not queue behind a large upload on the copy engine."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if device is None or torch.device(device).type != "cuda":
        return t if device is None else t.to(device)
    from . import _engine
    return _engine.to_device(np.ascontiguousarray(arr), t.dtype)


def all_gather_counts(counts, group=None, device=None):
    """numGenes of every rank's rows, rank-major (host int64 array)."""
    world_size, rank = world()
    counts = np.asarray(counts, dtype=np.int64)
    if world_size == 1:
        return counts
    t = _small_to(counts.reshape(1, -1).copy(), device)
    return all_gather_rows(t, group=group)[0].cpu().numpy()


class GlobalBackground(object):
    """Index bookkeeping for p-values over the gathered medians.

    Every rank's block is laid out [real rows | random rows]; the gathered
    matrix is rank-major.  ``real_cols`` / ``rnd_cols`` select the real and the
    random signatures of all ranks in (rank, row) order -- the order the
    reference would see had the caller concatenated the per-rank dictionaries.
    The numGenes grouping is that of Signatures.py:417-439."""

    def __init__(self, counts_all, n_real_per_rank, rows_per_rank, device):
        from .Signatures import _background_groups
        counts_all = np.asarray(counts_all, dtype=np.int64)
        world_size = len(rows_per_rank)
        real_cols, rnd_cols = [], []
        off = 0
        for r in range(world_size):
            real_cols.extend(range(off, off + n_real_per_rank[r]))
            rnd_cols.extend(range(off + n_real_per_rank[r], off + rows_per_rank[r]))
            off += rows_per_rank[r]
        self.real_cols_host = np.asarray(real_cols, dtype=np.int64)
        self.rnd_cols_host = np.asarray(rnd_cols, dtype=np.int64)
        order, gptr, grp = _background_groups(list(counts_all[self.rnd_cols_host]),
                                              list(counts_all[self.real_cols_host]))
        self.order_host, self.gptr_host, self.grp_host = order, gptr, grp
        self.device = device
        if device is not None and torch.device(device).type == "cuda":
            self.real_cols = _small_to(self.real_cols_host, device)
            self.rnd_cols = _small_to(self.rnd_cols_host, device)
            self.order = _small_to(order, device)
            self.gptr = _small_to(gptr, device)
            self.grp = _small_to(grp, device)

    def pvalues(self, all_med):
        """all_med: [P x rows_total] CUDA float64 (gathered) -> [P x K_total]
        p-values of every rank's real signatures, via fp_background_pvalues."""
        from . import _engine
        rows = []
        for i in range(all_med.shape[0]):
            med = all_med[i].index_select(0, self.real_cols).contiguous()
            rnd = all_med[i].index_select(0, self.rnd_cols).contiguous()
            p, _, _ = _engine.background_pvalues(med, self.grp, rnd, self.order, self.gptr)
            rows.append(p)
        return torch.stack(rows)
