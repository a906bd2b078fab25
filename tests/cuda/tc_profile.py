"""Profile driver: one warm-up + N timed launches of the tensor-core consistency
kernel on a C2-shaped slice (20 000 cells, K signatures).  Used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 635
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
any_order = not (len(sys.argv) > 4 and sys.argv[4] == "ordered")   # the product path only takes row medians
proj, _ = synth.make_projections(n, 1, seed=5)
c = torch.from_numpy(proj["Proj00"]).cuda()
ld = E.padded(n)
g = torch.Generator(device="cuda"); g.manual_seed(1)
scores = torch.rand((k, ld), dtype=torch.float64, device="cuda", generator=g)
ranks = E.rank_rows(scores, n)
out = torch.empty_like(ranks)
E.consistency(c, 0.33 ** 2, ranks, n, out=out, precision="tc", any_cell_order=any_order)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    E.consistency(c, 0.33 ** 2, ranks, n, out=out, precision="tc", any_cell_order=any_order)
    ev[r + 1].record()
torch.cuda.synchronize()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
flops = 2.0 * n * n * (k + 1)
print("n=%d k=%d: %s ms  -> %.1f TFLOP/s algorithmic" % (n, k, ["%.3f" % m for m in ms], flops / (min(ms) * 1e-3) / 1e12))
