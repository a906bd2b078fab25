"""world_size-2 gloo tests of the signature-block partitioning (CPU only).

The data path of the multi-GPU run has no collective; what is exchanged is the
[projections x rows] matrix of medians.  These tests run the exchange and the
index bookkeeping on two gloo ranks and compare the pooled background table
with the oracle's (Signatures.py:417-439) on the same medians."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastproject_b200 import distributed as D
from oracle import fp_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_block(rank):
    """Deterministic per-rank block: [real | random] medians for 2 projections."""
    rng = np.random.RandomState(100 + rank)
    n_real = 5 + rank                   # unequal widths on purpose
    sizes = [5, 10, 20]
    rnd_counts = np.repeat(sizes, 7 + rank)
    rng.shuffle(rnd_counts)
    real_counts = rng.randint(3, 40, size=n_real)
    counts = np.concatenate([real_counts, rnd_counts]).astype(np.int64)
    med = rng.normal(loc=300.0, scale=5.0, size=(2, counts.size))
    return n_real, counts, med


def _worker(rank, world_size, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        n_real, counts, med = _rank_block(rank)
        shape_all = D.all_gather_counts([n_real, counts.size]).reshape(world_size, 2)
        counts_all = D.all_gather_counts(counts)
        all_med = D.all_gather_rows(torch.from_numpy(med))
        bg = D.GlobalBackground(counts_all, [int(v) for v in shape_all[:, 0]],
                                [int(v) for v in shape_all[:, 1]], device=None)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), shape_all=shape_all,
                 counts_all=counts_all, all_med=all_med.numpy(), real_cols=bg.real_cols_host,
                 rnd_cols=bg.rnd_cols_host, order=bg.order_host, gptr=bg.gptr_host, grp=bg.grp_host)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 3500, 19001):
        for w in (1, 2, 3, 8):
            blocks = [D.shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks[:-1], blocks[1:]):
                assert a[1] == b[0]
            widths = [hi - lo for lo, hi in blocks]
            assert max(widths) - min(widths) <= 1


@pytest.mark.timeout(120)
def test_two_rank_median_exchange_and_background(tmp_path):
    world_size = 2
    mp.spawn(_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
    got = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(world_size)]
    blocks = [_rank_block(r) for r in range(world_size)]
    want_med = np.concatenate([b[2] for b in blocks], axis=1)
    want_counts = np.concatenate([b[1] for b in blocks])
    for g in got:                                   # every rank sees the same, rank-major
        assert np.array_equal(g["all_med"], want_med)
        assert np.array_equal(g["counts_all"], want_counts)
    g = got[0]
    # the pooled background equals the oracle's table on the concatenated dictionaries
    real_cols, rnd_cols = g["real_cols"], g["rnd_cols"]
    assert len(real_cols) == sum(b[0] for b in blocks)
    for proj in range(want_med.shape[0]):
        rnd_med = want_med[proj, rnd_cols]
        rnd_ng = want_counts[rnd_cols]
        table = O.background_table(rnd_med, rnd_ng)
        # groups in first-seen order, members in original order
        for grp in range(table.shape[0]):
            members = g["order"][g["gptr"][grp]:g["gptr"][grp + 1]]
            assert np.all(rnd_ng[members] == table[grp, 0])
            assert np.array_equal(np.sort(members), members)
            assert np.isclose(np.mean(rnd_med[members]), table[grp, 1], rtol=1e-15)
        # nearest-numGenes assignment of every real signature (argmin, first wins)
        for k, col in enumerate(real_cols):
            want = int(np.argmin(np.abs(want_counts[col] - table[:, 0])))
            assert g["grp"][k] == want
