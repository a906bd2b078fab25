"""Development check: tensor-core consistency (FP_PRECISION_TC) against the FP64
verifier kernel on the same inputs.  Prints error statistics; run on a B200."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, _native, synth
from scipy.stats import rankdata


def case(n, k, seed, outliers=0, dup=0, per_cell=False, kind=_native.OUT_ABS_ERROR, binary=False):
    rng = np.random.RandomState(seed)
    proj, _ = synth.make_projections(n, 1, seed=seed)
    coords = proj["Proj00"].copy()
    for o in range(outliers):
        coords[:, o] = [6.0 + 3 * o, -5.0 - o]
    for d in range(dup):
        coords[:, n - 1 - d] = coords[:, d + 5]
    if binary:
        vals = (rng.uniform(size=(k, n)) < 0.3).astype(np.float64)
    else:
        vals = np.array([rankdata(rng.normal(size=n) + (coords[0] if r % 3 == 0 else 0)) for r in range(k)])
        if k > 2:
            vals[1] = rankdata(np.round(rng.normal(size=n), 1))      # many ties -> half-integers
    ld = E.padded(n)
    dev = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(vals).cuda()
    c = torch.from_numpy(coords).cuda()
    sig = 0.33 ** 2
    if per_cell:
        sig = torch.from_numpy((0.2 + 0.3 * rng.uniform(size=n)) ** 2).cuda()
    a = E.consistency(c, sig, dev, n, precision="fp64", output_kind=kind)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = E.consistency(c, sig, dev, n, precision="tc", output_kind=kind)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    a = a[:, :n].cpu().numpy(); b = b[:, :n].cpu().numpy()
    diff = np.abs(a - b)
    worst = np.unravel_index(np.argmax(diff), diff.shape)
    ma, mb = np.median(a, axis=1), np.median(b, axis=1)
    print("n=%d k=%d outl=%d dup=%d percell=%d kind=%d bin=%d: max|d|=%.3e (rel N %.2e) at %s  mean|d|=%.3e  "
          "max|dmedian|=%.3e  nan=%d  tc %.2f ms" % (n, k, outliers, dup, per_cell, kind, binary, diff.max(),
          diff.max() / n, worst, diff.mean(), np.abs(ma - mb).max(), int(np.isnan(b).sum()), dt * 1e3))
    return diff.max() / n


worst = 0.0
worst = max(worst, case(300, 3, 1))
worst = max(worst, case(1000, 5, 2))
worst = max(worst, case(4097, 130, 3, outliers=2, dup=3))
worst = max(worst, case(2000, 40, 4, per_cell=True))
worst = max(worst, case(2000, 300, 5, kind=_native.OUT_PREDICTION, binary=True))
worst = max(worst, case(20000, 260, 6, outliers=1))
worst = max(worst, case(33000, 130, 7, outliers=1))
worst = max(worst, case(40000, 20, 8))
print("WORST rel", worst)
