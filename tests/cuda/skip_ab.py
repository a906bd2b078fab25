"""A/B on one box: the tensor-core contraction with the cells in Z-order (all-zero weight digit planes
skipped) against the given cell order (nothing to skip), same inputs, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, synth

def run(n, k, n_proj, reps):
    proj, _ = synth.make_projections(n, n_proj, seed=1335607829)
    names = sorted(proj)
    ld = E.padded(n)
    vals = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    base = torch.argsort(torch.rand((64, n), device="cuda", dtype=torch.float64), dim=1).to(torch.float64) + 1.0
    for lo in range(0, k, 64):
        hi = min(k, lo + 64)
        vals[lo:hi, :n] = base[:hi - lo].roll(lo, dims=1)
    out = torch.empty_like(vals)
    coords = [torch.from_numpy(proj[p]).cuda() for p in names]
    res = {}
    for mode in ("zorder", "given"):
        if mode == "given":
            os.environ["FASTPROJECT_B200_CELL_ORDER"] = "given"
        else:
            os.environ.pop("FASTPROJECT_B200_CELL_ORDER", None)
        for c in coords:
            E.consistency(c, 0.33 ** 2, vals, n, out=out, precision="tc", any_cell_order=True, slot="ab")
        torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            for c in coords:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                E.consistency(c, 0.33 ** 2, vals, n, out=out, precision="tc", any_cell_order=True, slot="ab")
                e1.record()
                ts.append((e0, e1))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ts]
        per = np.array(ms).reshape(reps, len(coords)).mean(axis=0)
        res[mode] = per
        print("%-7s n=%d k=%d: per projection ms %s  mean %.3f" % (mode, n, k, np.round(per, 3), per.mean()), flush=True)
    print("ratio zorder/given per projection:", np.round(res["zorder"] / res["given"], 3), "mean %.3f" % (res["zorder"].mean() / res["given"].mean()))

run(20000, 3500, 6, 5)
run(100000, 4000, 2, 1)
