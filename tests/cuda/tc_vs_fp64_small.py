"""Small tensor-core vs FP64 check (used under compute-sanitizer memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, synth
from scipy.stats import rankdata
n, k = 700, 130
rng = np.random.RandomState(1)
proj, _ = synth.make_projections(n, 1, seed=3)
vals = np.array([rankdata(rng.normal(size=n)) for _ in range(k)])
ld = E.padded(n)
dev = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
dev[:, :n] = torch.from_numpy(vals).cuda()
c = torch.from_numpy(proj["Proj00"]).cuda()
a = E.consistency(c, 0.33 ** 2, dev, n, precision="fp64")[:, :n].cpu().numpy()
b = E.consistency(c, 0.33 ** 2, dev, n, precision="tc")[:, :n].cpu().numpy()
med = E.row_medians(E.consistency(c, 0.33 ** 2, dev, n, precision="tc"), n).cpu().numpy()
print("max diff", np.abs(a - b).max(), "median[0]", med[0])
