"""Profile driver for fp_score_signatures on a C2-shaped input (used under ncu)."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, _scoring, Signatures as S

G = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
per_size = int(sys.argv[3]) if len(sys.argv) > 3 else 500
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x = torch.randn((G, N), dtype=torch.float64, device="cuda")
w = torch.rand((G, N), dtype=torch.float64, device="cuda")
genes = ["G%05d" % i for i in range(G)]
random.seed(1); np.random.seed(1)
sigs = S.generate_random_sigs(genes, False, [5, 10, 20, 50, 100, 200], per_size)
kept, sp, sg, ss, counts = _scoring.pack_signatures(sigs, genes, 5)
dp, dg, ds = E.to_device(sp, torch.int32), E.to_device(sg, torch.int32), E.to_device(ss, torch.int8)
E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
    ev[r + 1].record()
torch.cuda.synchronize()
ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
nnz = int(sp[-1])
print("G=%d N=%d K=%d nnz=%d: %.3f ms -> %.2f G scores/s; compulsory %.1f GB/s, gather %.1f GB/s" % (
    G, N, len(kept), nnz, ms, N * len(kept) / ms / 1e6, (G * N * 16 + N * len(kept) * 8) / ms / 1e6,
    16.0 * nnz * N / ms / 1e6))
