"""Timing of the normalisation kernels at the C2 shape (15 000 genes x 20 000 cells)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import _engine as E
G, N = 15000, 20000
x = torch.randn((G, N), dtype=torch.float64, device="cuda")
for name, fn in (("rows", E.normalize_rows), ("cols", E.normalize_cols)):
    fn(x); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for r in range(2):
        fn(x); ev[r + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(2))
    print("normalize_%s %dx%d: %.2f ms  (%.0f GB/s of 3 reads + 1 write)" % (name, G, N, ms, G * N * 8 * 4 / ms / 1e6))
xh = x[:2000].cpu().numpy()
t0 = time.perf_counter()
mu = xh.mean(axis=1, keepdims=True); sd = xh.std(axis=1, keepdims=True); sd[sd == 0] = 1; y = (xh - mu) / sd
print("numpy row_normalization on 2000 of the rows: %.1f ms -> %.0f ms for all" % ((time.perf_counter() - t0) * 1e3, (time.perf_counter() - t0) * 1e3 * G / 2000))
# false-negative weights (fp_compute_weights) at the same shape
xe = torch.where(torch.rand((G, N), device="cuda") < 0.5, torch.zeros_like(x), x.abs())
params = torch.stack([torch.rand(N, dtype=torch.float64, device="cuda") * 3 + 0.5,
                      torch.rand(N, dtype=torch.float64, device="cuda") * 2,
                      torch.zeros(N, dtype=torch.float64, device="cuda"),
                      torch.ones(N, dtype=torch.float64, device="cuda")])
E.compute_weights(xe, params); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for r in range(2):
    E.compute_weights(xe, params); ev[r + 1].record()
torch.cuda.synchronize()
ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(2))
print("compute_weights %dx%d: %.2f ms  (%.0f GB/s of 1 read + 1 write)" % (G, N, ms, G * N * 8 * 2 / ms / 1e6))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fp_oracle as O
xh = xe[:300].cpu().numpy(); ph = params.cpu().numpy()
t0 = time.perf_counter(); O.compute_weights(O.logistic_fnr, ph, xh, O.estimate_non_detection(xh)); dt = time.perf_counter() - t0
print("numpy compute_weights on 300 of the rows: %.0f ms -> %.1f s for all" % (dt * 1e3, dt * G / 300))
