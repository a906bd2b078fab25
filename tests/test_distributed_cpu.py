"""world_size-2/3 gloo tests of the multi-GPU partitioning (CPU only).

One job, split two ways (fastproject_b200/distributed.py): rows of the signature batch for
scoring + ranking, then runs of the (projection x signature-block) grid for the contraction
and the medians.  The exchanges are one all-gather of row blocks and one all-reduce of the
[projections x rows] medians.  These tests run both on gloo ranks with a CPU stand-in for the
per-run kernel and compare with the same job done by one process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastproject_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 3500, 19001):
        for w in (1, 2, 3, 8):
            blocks = [D.shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks[:-1], blocks[1:]):
                assert a[1] == b[0]
            widths = [hi - lo for lo, hi in blocks]
            assert max(widths) - min(widths) <= 1


def test_cyclic_owner_rows_cover_and_balance():
    for n, w in ((0, 4), (1, 8), (7, 8), (3500, 8), (3500, 2), (19001, 3)):
        rows = D.cyclic_owner_rows(n, w)
        assert len(rows) == w
        allr = np.sort(np.concatenate(rows)) if n else np.zeros(0, dtype=np.int64)
        assert np.array_equal(allr, np.arange(n))
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1
        for r in rows:
            assert np.all(np.diff(r) > 0)
    # the job's order: 500 real signatures then the background by size -- contiguous blocks would
    # give the last rank 5x the genes of the lightest; round-robin gives every rank the same mix
    genes = np.concatenate([np.full(500, 80), np.repeat([5, 10, 20, 50, 100, 200], 500)])
    sums = [genes[r].sum() for r in D.cyclic_owner_rows(genes.size, 8)]
    assert max(sums) / min(sums) < 1.02
    even = [genes[lo:hi].sum() for lo, hi in D.even_bounds(genes.size, 8)]
    assert max(even) / min(even) > 4
    # with rejected signatures the rows follow the POSITIONS in the caller's list
    positions = np.array([0, 1, 3, 4, 8, 9, 10, 15])
    rows = D.cyclic_owner_rows(positions.size, 4, positions=positions)
    assert [list(r) for r in rows] == [[0, 3, 4], [1, 5], [6], [2, 7]]


@pytest.mark.parametrize("n_proj,n_rows,world", [
    (6, 3500, 8), (8, 4000, 8), (8, 19000, 8), (4, 3200, 8), (6, 3500, 1), (6, 3500, 2), (6, 3500, 4),
    (1, 10, 8), (3, 254, 2), (2, 255, 3), (5, 0, 4), (0, 100, 2)])
def test_partition_units_cover_every_pair_once(n_proj, n_rows, world):
    runs = D.partition_units(n_proj, n_rows, world)
    assert len(runs) == world
    owner = np.zeros((n_proj, n_rows), dtype=np.int64)
    for r, mine in enumerate(runs):
        for p, lo, hi in mine:
            assert 0 <= p < n_proj and 0 <= lo < hi <= n_rows
            assert lo % D.SIG_BLOCK == 0                       # runs start on a CTA-pair row tile
            owner[p, lo:hi] += 1
        # one call of the contraction per projection touched
        assert len(set(p for p, _, _ in mine)) == len(mine)
    assert (owner == 1).all()
    # every unit is dealt, and the cost (units + RUN_OVERHEAD per projection touched) is balanced:
    # no rank carries more than the ideal share plus one unit and one projection overhead
    n_blocks = (n_rows + D.SIG_BLOCK - 1) // D.SIG_BLOCK
    units = [sum((hi - lo + D.SIG_BLOCK - 1) // D.SIG_BLOCK for _, lo, hi in mine) for mine in runs]
    assert sum(units) == n_proj * n_blocks
    cost = [u + D.RUN_OVERHEAD * len(mine) for u, mine in zip(units, runs)]
    if n_proj * n_blocks >= world:
        assert max(cost) <= sum(cost) / world + 1 + D.RUN_OVERHEAD + 1e-9


def test_partition_units_is_projection_major_when_projections_divide():
    # C3: 8 projections on 8 GPUs -> one whole projection per rank (its weight planes are
    # generated once in the job); on 4 and 2 GPUs whole projections too
    for world in (8, 4, 2):
        runs = D.partition_units(8, 4000, world)
        for r, mine in enumerate(runs):
            assert [(lo, hi) for _, lo, hi in mine] == [(0, 4000)] * (8 // world)
    # C2: 6 projections on 8 GPUs -> a rank straddles at most two projections
    for mine in D.partition_units(6, 3500, 8):
        assert 1 <= len(mine) <= 2


def _job(seed=5, n_rows=37, n_cells=50, n_proj=3):
    rng = np.random.RandomState(seed)
    scores = rng.normal(size=(n_rows, n_cells))
    shifts = rng.normal(size=(n_proj,))
    return scores, shifts


def _stand_in_rank(x):
    return np.argsort(np.argsort(x, axis=1), axis=1).astype(np.float64) + 1.0


def _stand_in_median(ranks_block, shift):
    return np.median(np.abs(ranks_block - ranks_block.mean(axis=1, keepdims=True) * shift), axis=1)


def _worker(rank, world_size, port, out_dir, bounds_kind):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        scores, shifts = _job()
        n_rows, n_proj = scores.shape[0], shifts.size
        if bounds_kind == "even":
            bounds = D.even_bounds(n_rows, world_size)
        else:                                              # what rejected signatures leave behind
            cuts = [0, 3, 3 + 20][:world_size] + [n_rows]
            bounds = [(cuts[r], cuts[r + 1]) for r in range(world_size)]
        lo, hi = bounds[rank]
        # stage 1: rank my rows, gather all -- contiguous blocks (how the expression rows travel) ...
        local = torch.from_numpy(_stand_in_rank(scores[lo:hi]))
        ranks = D.all_gather_row_blocks(local, bounds)
        # ... and round-robin rows (how the signature rows are dealt): same matrix, job order
        if bounds_kind == "even":
            owner = D.cyclic_owner_rows(n_rows, world_size)
        else:
            owner = D.cyclic_owner_rows(n_rows, world_size, positions=np.arange(n_rows) * 3 + (np.arange(n_rows) % 2))
        mine = torch.from_numpy(_stand_in_rank(scores[owner[rank]]))
        assert torch.equal(D.all_gather_owned_rows(mine, owner), ranks)
        # average ranks travel as float32 (exact for integers and halves) and arrive as float64
        halves = mine + 0.5 * (mine % 2)
        got32 = D.all_gather_rank_rows(halves, owner, scores.shape[1])
        assert got32.dtype == torch.float64
        assert torch.equal(got32, D.all_gather_owned_rows(halves, owner))
        # stage 2: my runs of the (projection x block) grid (tiny blocks so that runs straddle)
        med = torch.zeros((n_proj, n_rows), dtype=torch.float64)
        for p, rlo, rhi in D.partition_units(n_proj, n_rows, world_size, block=5)[rank]:
            med[p, rlo:rhi] = torch.from_numpy(_stand_in_median(ranks[rlo:rhi].numpy(), shifts[p]))
        D.all_reduce_sum(med)
        got = D.broadcast_object({"from": rank} if rank == 0 else None)
        assert got == {"from": 0}
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ranks=ranks.numpy(), med=med.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world_size,bounds_kind", [(2, "even"), (3, "ragged")])
def test_gloo_ranks_reproduce_the_single_process_job(tmp_path, world_size, bounds_kind):
    mp.spawn(_worker, args=(world_size, _free_port(), str(tmp_path), bounds_kind), nprocs=world_size,
             join=True)
    scores, shifts = _job()
    want_ranks = _stand_in_rank(scores)
    want_med = np.stack([_stand_in_median(want_ranks, s) for s in shifts])
    for r in range(world_size):                    # every rank ends with the whole result, bit for bit
        got = np.load(str(tmp_path / ("rank%d.npz" % r)))
        assert np.array_equal(got["ranks"], want_ranks)
        assert np.array_equal(got["med"], want_med)


def test_single_process_helpers_are_identity():
    t = torch.arange(12, dtype=torch.float64).reshape(3, 4)
    assert D.world() == (1, 0)
    assert D.all_gather_row_blocks(t, [(0, 3)]) is t
    assert D.all_reduce_sum(t) is t
    assert D.broadcast_object("x") == "x"
