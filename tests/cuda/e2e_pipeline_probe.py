"""Per-step timing of the pipelined end-to-end loop (prefetch of the next step's matrices)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import __graft_entry__ as ge
ge.build()
import fastproject_b200 as FP
from fastproject_b200 import Signatures as S, _engine as E, synth
cfg = bench.WORKLOADS["C2"]
wl = bench.make_workload(cfg, torch.device("cuda", 0))
G, N = cfg["G"], cfg["N"]
host = []
for _ in range(2):
    xh = torch.empty((G, N), dtype=torch.float64, pin_memory=True); wh = torch.empty((G, N), dtype=torch.float64, pin_memory=True)
    xh.copy_(wl["x"]); wh.copy_(wl["w"])
    host.append(synth.ExpressionMatrix(xh.numpy(), wh.numpy(), wl["genes"], wl["cells"]))
torch.cuda.synchronize()
def step(data, nxt, log):
    t = [time.perf_counter()]
    if nxt is not None:
        FP.prefetch(nxt.base, nxt.weights)
    real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5); t.append(time.perf_counter())
    t.append(time.perf_counter())
    rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5); t.append(time.perf_counter())
    res = S.sigs_vs_projections(wl["projections"], real, rnd, 0.33); t.append(time.perf_counter())
    hs = next(iter(real.values()))._batch.host_scores(); t.append(time.perf_counter())
    E.expression_cache.evict(data.base); E.expression_cache.evict(data.weights)
    log.append(["%.1f" % (1e3 * (b - a)) for a, b in zip(t[:-1], t[1:])])
log = []
E.clear_device_cache(); step(host[0], None, log); E.clear_device_cache(); torch.cuda.synchronize()
t0 = time.perf_counter(); marks = []
for it in range(6):
    step(host[it % 2], host[(it + 1) % 2] if it < 5 else None, log)
    marks.append(time.perf_counter() - t0)
    st = torch.cuda.memory_stats()
    print("   step %d: cudaMalloc calls so far %d, cudaFree %d, reserved %.1f GB" % (it, st["num_device_alloc"], st["num_device_free"], st["reserved_bytes.all.current"] / 1e9))
torch.cuda.synchronize()
print("cumulative ms per step:", ["%.0f" % (m * 1e3) for m in marks], "total %.0f" % ((time.perf_counter() - t0) * 1e3))
for l in log: print("  [score real, prefetch call, score rnd, svp, d2h]:", l)
