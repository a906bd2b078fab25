# -*- coding: utf-8 -*-
"""Multi-GPU partitioning of the scoring core: one process per GPU, ONE job.

Every rank is handed the same inputs (expression data, the full signature lists, all
projections) and returns the same result as a single-GPU run; the work of the job is split:

  stage 1  scoring + ranking : the signatures (real and random background alike) are dealt
           round-robin to the ranks (``cyclic_owner_rows``): the same number of rows and the same
           mix of signature sizes everywhere.  A rank reads X / W once and writes only its rows.
  exchange ``all_gather_owned_rows``: the rank rows (float64 [rows x cells]) are all-gathered
           over NVLink and put back in job order, so that every rank can contract any block of
           signatures against any projection (C2: 560 MB, C3 with 4 000 signatures: 3.2 GB).
  stage 2  consistency + medians : the (projection x signature-block) grid -- blocks of 254
           signatures, the row tile of one CTA pair of the tensor-core kernel -- is laid out
           projection-major and cut into ``world`` contiguous runs (``partition_units``).  When
           the number of projections is a multiple of the number of GPUs (C3: 8 / 8) a rank
           owns whole projections and generates each projection's weight planes exactly once
           in the job; otherwise (C2: 6 projections on 8 GPUs) a run straddles at most two
           projections.
  gather   ``all_reduce_sum`` of the [projections x signatures] medians, zero outside a rank's
           own units (every entry has exactly one owner, so the sum is exact) -- C3: 256 KB.
           The background N(mu, sigma) of a numGenes group pools random signatures of every
           block (/root/reference/FastProject/Signatures.py:417-430), so every rank needs all
           medians; p-values are then computed on every rank.

The reference has no counterpart (it is single-process).  The helpers work on any
``torch.distributed`` backend: NCCL on the GPUs; gloo in the CPU tests and in the
two-ranks-on-one-GPU parity test, where CUDA tensors are staged through the host.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

SIG_BLOCK = 254          # signatures per CTA pair of consistency_mma_kernel (2 x 127)


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


def cyclic_owner_rows(n_rows, world_size, positions=None):
    """Which rows of the job's signature batch each rank scores and ranks: dealt round-robin by
    POSITION in the caller's signature list (``positions[j]`` = list position of job row j;
    default: j itself).  Contiguous blocks would hand one rank all the 200-gene background
    signatures and another all the 5-gene ones (5x the scoring work, and everybody waits for the
    slowest at the all-gather); round-robin gives every rank the same mix and the same number of
    rows.  Returns a list of ascending int64 index arrays, one per rank."""
    pos = np.arange(int(n_rows), dtype=np.int64) if positions is None else np.asarray(positions, dtype=np.int64)
    return [np.nonzero(pos % world_size == r)[0] for r in range(int(world_size))]


def all_gather_owned_rows(local, owner_rows, group=None):
    """local: this rank's rows (``owner_rows[rank]``, in that order) of a [n_rows x ld] matrix ->
    the whole matrix in job order on every rank: one all-gather of blocks padded to the widest,
    then one gather-copy that puts the rows back in order."""
    world_size, rank = world()
    if world_size == 1:
        return local
    ld = local.shape[1]
    widest = max(len(r) for r in owner_rows)
    n_rows = sum(len(r) for r in owner_rows)
    mine = local
    if local.shape[0] != widest:
        mine = torch.zeros((widest, ld), dtype=local.dtype, device=local.device)
        mine[:local.shape[0]] = local
    buf = torch.empty((world_size * widest, ld), dtype=local.dtype, device=local.device)
    all_gather_into(buf, mine.contiguous(), group)
    src = np.empty(n_rows, dtype=np.int64)
    for r, rows in enumerate(owner_rows):
        src[rows] = r * widest + np.arange(len(rows))
    return buf.index_select(0, _index_to(src, local.device))


def all_gather_rank_rows(local_ranks, owner_rows, n_cells, group=None):
    """``all_gather_owned_rows`` for average ranks: they are integers or half-integers up to
    ``n_cells``, exactly representable in float32 below 2^22 cells, so half the bytes cross NVLink
    and the rows are widened back to float64 on arrival (bit-identical)."""
    world_size, _ = world()
    if world_size == 1:
        return local_ranks
    if int(n_cells) < (1 << 22) and local_ranks.dtype == torch.float64:
        return all_gather_owned_rows(local_ranks.to(torch.float32), owner_rows, group).to(torch.float64)
    return all_gather_owned_rows(local_ranks, owner_rows, group)


def _index_to(arr, device):
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if torch.device(device).type != "cuda":
        return t
    from . import _engine
    return _engine.to_device(np.ascontiguousarray(arr), torch.long)


RUN_OVERHEAD = 2.5       # what a rank pays per projection it touches, in units of one signature block:
                         # weight planes + value packing + cell order, ~1.0 ms against 0.39 ms per
                         # block at 20 000 cells (measured, C2 on 8 GPUs)


def partition_units(n_proj, n_rows, world_size, block=SIG_BLOCK, run_overhead=RUN_OVERHEAD):
    """(projection x signature-block) units, projection-major, cut into ``world_size``
    contiguous runs of (nearly) equal COST: a unit costs 1, and every projection a rank touches
    costs it ``run_overhead`` more (its weight planes are generated per rank) -- so a rank whose
    run straddles two projections gets fewer units than one that stays inside a projection.  The
    cut minimises the largest cost (bisection on the bound, greedy feasibility).  Returns, per
    rank, a list of ``(projection, row_lo, row_hi)`` -- the units of one projection merged into one
    row range, so a rank calls the contraction once per projection it touches."""
    n_proj, n_rows, world_size = int(n_proj), int(n_rows), int(world_size)
    n_blocks = (n_rows + block - 1) // block
    total = n_proj * n_blocks

    def cuts_for(bound):
        """Greedy: give each rank as many consecutive units as fit under ``bound``."""
        cuts, u = [0], 0
        for _ in range(world_size):
            cost, start = 0.0, u
            while u < total:
                first_of_projection = (u == start) or (u % n_blocks == 0)
                step = 1.0 + (run_overhead if first_of_projection else 0.0)
                if cost + step > bound + 1e-9 and u > start:
                    break
                cost += step
                u += 1
            cuts.append(u)
        return cuts if u == total else None

    if total == 0:
        return [[] for _ in range(world_size)]
    lo_b, hi_b = 0.0, float(total) + run_overhead * (n_proj + world_size)
    for _ in range(50):
        mid = 0.5 * (lo_b + hi_b)
        if cuts_for(mid) is None:
            lo_b = mid
        else:
            hi_b = mid
    cuts = cuts_for(hi_b)
    out = []
    for r in range(world_size):
        lo, hi = cuts[r], cuts[r + 1]
        runs = []
        u = lo
        while u < hi:
            p = u // n_blocks
            b_lo = u - p * n_blocks
            b_hi = min(n_blocks, b_lo + (hi - u))
            runs.append((p, b_lo * block, min(n_rows, b_hi * block)))
            u += b_hi - b_lo
        out.append(runs)
    return out


def _is_gloo_cuda(t, group=None):
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce_sum(t, group=None):
    """In-place sum over the ranks (no-op on one rank)."""
    world_size, _ = world()
    if world_size == 1:
        return t
    if _is_gloo_cuda(t, group):
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_into(buf, mine, group=None):
    """``dist.all_gather_into_tensor`` that also works for CUDA tensors on a gloo group (the
    two-ranks-on-one-GPU test): staged through the host there."""
    if _is_gloo_cuda(mine, group):
        hb = torch.empty(buf.shape, dtype=buf.dtype)
        dist.all_gather_into_tensor(hb, mine.cpu(), group=group)
        buf.copy_(hb)
    else:
        dist.all_gather_into_tensor(buf, mine, group=group)
    return buf


_host_group = None


def host_group():
    """A gloo group over the same ranks, for small HOST-side exchanges that must not wait for
    (or be ordered behind) the work queued on the CUDA stream.  Created on first use: every rank
    must reach that first use (``calculate_sig_scores(distributed=True)`` does)."""
    global _host_group
    if _host_group is None:
        _host_group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    return _host_group


def all_reduce_sum_host(arr):
    """Sum of a small NumPy array over the ranks through the host group (no GPU involvement)."""
    world_size, _ = world()
    if world_size == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=host_group())
    return t.numpy()


def even_bounds(n_rows, world_size):
    return [shard_bounds(n_rows, world_size, r) for r in range(world_size)]


def all_gather_row_blocks(local, bounds, group=None):
    """local: this rank's block ``bounds[rank] = (lo, hi)`` of a [n_rows x ld] matrix -> the
    whole matrix on every rank (one all-gather; the blocks are padded to the widest for the
    collective and compacted afterwards if they are not all equal)."""
    world_size, rank = world()
    if world_size == 1:
        return local
    ld = local.shape[1]
    widths = [hi - lo for lo, hi in bounds]
    widest = max(widths)
    mine = local
    if local.shape[0] != widest:
        mine = torch.zeros((widest, ld), dtype=local.dtype, device=local.device)
        mine[:local.shape[0]] = local
    mine = mine.contiguous()
    buf = torch.empty((world_size * widest, ld), dtype=local.dtype, device=local.device)
    all_gather_into(buf, mine, group)
    if all(w == widest for w in widths):
        return buf
    return torch.cat([buf[r * widest:r * widest + widths[r]] for r in range(world_size)], dim=0)


def broadcast_object(obj, src=0, group=None):
    world_size, _ = world()
    if world_size == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]
