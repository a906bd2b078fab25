"""Does a concurrent host->device copy slow the kernels?  Times each kernel of the path with
CUDA events, alone and while a 2.4 GB pinned upload runs on a side stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E, synth
n, k = 20000, 3500
proj, _ = synth.make_projections(n, 1, seed=5)
c = torch.from_numpy(proj["Proj00"]).cuda()
ld = E.padded(n)
scores = torch.rand((k, ld), dtype=torch.float64, device="cuda")
ranks = E.rank_rows(scores, n)
out = torch.empty_like(ranks)
host = torch.empty((15000, 20000), dtype=torch.float64, pin_memory=True); host.fill_(1.0)
dst = torch.empty((15000, 20000), dtype=torch.float64, device="cuda")
side = torch.cuda.Stream()
def timed(fn, with_copy):
    torch.cuda.synchronize()
    if with_copy:
        with torch.cuda.stream(side):
            dst.copy_(host, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for name, fn in (("consistency tc", lambda: E.consistency(c, 0.33 ** 2, ranks, n, out=out, precision="tc")),
                 ("rank", lambda: E.rank_rows(scores, n)),
                 ("median", lambda: E.row_medians(out, n))):
    fn()
    a = min(timed(fn, False) for _ in range(2)); b = min(timed(fn, True) for _ in range(2))
    print("%-16s alone %.2f ms, with concurrent H2D copy %.2f ms" % (name, a, b))
