import sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import __graft_entry__ as ge; ge.build()
from fastproject_b200 import _engine as E
G, N = 15000, 20000
xh = torch.empty((G, N), dtype=torch.float64, pin_memory=True); xh.fill_(1.0)
a = xh.numpy()
h = torch.from_numpy(a)
print("pinned:", xh.is_pinned(), h.is_pinned())
torch.cuda.synchronize()
t0 = time.perf_counter(); E.expression_cache.prefetch(a); t1 = time.perf_counter()
print("prefetch call returned after %.1f ms" % ((t1 - t0) * 1e3))
torch.cuda.synchronize(); print("copy done after %.1f ms" % ((time.perf_counter() - t0) * 1e3))
# overlap test: big kernel on compute stream + prefetch
y = torch.randn((8192, 8192), device="cuda")
E.clear_device_cache(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40): z = y @ y
E.expression_cache.prefetch(a)
for _ in range(40): z = y @ y
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("enqueue %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
E.clear_device_cache(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(80): z = y @ y
torch.cuda.synchronize(); print("matmuls alone %.1f ms" % ((time.perf_counter() - t0) * 1e3))
