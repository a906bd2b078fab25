"""One tensor-core contraction at BASELINE config 3 shape (100 000 cells x 4 000 rank rows), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
proj, _ = synth.make_projections(n, 1, seed=1335607829)
ld = E.padded(n)
vals = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
base = torch.argsort(torch.rand((64, n), device="cuda", dtype=torch.float64), dim=1).to(torch.float64) + 1.0
for lo in range(0, k, 64):
    hi = min(k, lo + 64)
    vals[lo:hi, :n] = base[:hi - lo].roll(lo, dims=1)
out = torch.empty_like(vals)
c = torch.from_numpy(proj["Proj00"]).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
E.consistency(c, 0.33 ** 2, vals, n, out=out, precision="tc", any_cell_order=True, slot="c3p")
e1.record()
torch.cuda.synchronize()
print("n=%d k=%d: %.2f ms" % (n, k, e0.elapsed_time(e1)))
