"""BASELINE config 5: signature-scoring-only sweep (cells x signatures), weighted_avg.
Prints one markdown row per shape: time, scores/s, compulsory and gather bandwidth."""
import sys, os, random, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, _scoring, Signatures as S

SHAPES = [  # genes, cells, signatures
    (15000, 10000, 1000), (15000, 10000, 10000), (15000, 10000, 50000),
    (15000, 100000, 1000), (15000, 100000, 10000), (15000, 100000, 50000),
    (2000, 1000000, 1000), (2000, 1000000, 10000),
]
print("| genes | cells | signatures | nnz | ms | G scores/s | compulsory GB/s (frac of 6540) | gather GB/s |")
print("|---|---|---|---|---|---|---|---|")
for G, N, K in SHAPES:
    gen = torch.Generator(device="cuda"); gen.manual_seed(G + N)
    x = torch.randn((G, N), dtype=torch.float64, device="cuda", generator=gen)
    w = torch.rand((G, N), dtype=torch.float64, device="cuda", generator=gen)
    genes = ["G%05d" % i for i in range(G)]
    random.seed(1); np.random.seed(1)
    sizes = [5, 10, 20, 50, 100, 200]
    sigs = S.generate_random_sigs(genes, False, sizes, (K + 5) // 6)[:K]
    kept, sp, sg, ss, counts = _scoring.pack_signatures(sigs, genes, 5)
    dp, dg, ds = E.to_device(sp, torch.int32), E.to_device(sg, torch.int32), E.to_device(ss, torch.int8)
    out = E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
    torch.cuda.synchronize()
    del out
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    outs = []
    for r in range(2):
        o = E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
        ev[r + 1].record()
        del o
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(2))
    nnz = int(sp[-1])
    comp = (G * N * 16 + N * len(kept) * 8 + 5 * nnz) / ms / 1e6
    print("| %d | %d | %d | %d | %.2f | %.2f | %.0f (%.3f) | %.0f |" % (
        G, N, len(kept), nnz, ms, N * len(kept) / ms / 1e6, comp, comp / 6539.9, 16.0 * nnz * N / ms / 1e6), flush=True)
    del x, w, dp, dg, ds
    gc.collect(); torch.cuda.empty_cache()
