"""Timing of fp_row_medians on dissimilarity-like rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E
for n, k in ((20000, 3500), (100000, 1000), (5000, 3500)):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    ld = E.padded(n)
    v = (torch.rand((k, ld), dtype=torch.float64, device="cuda", generator=g) * n * 0.5).abs()
    E.row_medians(v, n); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for r in range(3):
        m = E.row_medians(v, n); ev[r + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(3))
    ref = np.median(v[:64, :n].cpu().numpy(), axis=1)
    print("n=%d k=%d: %.3f ms (%.1f GB/s of one sweep) exact=%s" % (n, k, ms, k * n * 8 / ms / 1e6,
          bool(np.array_equal(ref, m[:64].cpu().numpy()))))
