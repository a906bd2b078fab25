"""Host -> device copy rate of pinned memory on this box, alone and with the step's kernels
running (the e2e floor of the C2 workload is 4.8 GB / this rate)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
G, N = 15000, 20000
xh = torch.empty((G, N), dtype=torch.float64, pin_memory=True); xh.fill_(1.0)
d = torch.empty((G, N), dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        d.copy_(xh, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("H2D alone: %.1f ms  %.1f GB/s" % (dt * 1e3, xh.nbytes / dt / 1e9))
y = torch.randn((8192, 8192), device="cuda", dtype=torch.float32)
big = torch.empty((1 << 28,), dtype=torch.float32, device="cuda")
for name, fn in (("matmul", lambda: y @ y), ("hbm-stream", lambda: big.add_(1.0))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(xh, non_blocking=True); e1.record()
    for _ in range(60): fn()
    torch.cuda.synchronize()
    print("H2D under %s: %.1f ms  %.1f GB/s" % (name, e0.elapsed_time(e1), xh.nbytes / e0.elapsed_time(e1) / 1e6))
