"""Where the end-to-end time of the public API goes (C2 shape, host inputs)."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import __graft_entry__ as ge
ge.build()
from fastproject_b200 import Signatures as S, _engine as E, synth
cfg = bench.WORKLOADS["C2"]
dev = torch.device("cuda", 0)
wl = bench.make_workload(cfg, dev)
G, N = cfg["G"], cfg["N"]
xh = torch.empty((G, N), dtype=torch.float64, pin_memory=True); wh = torch.empty((G, N), dtype=torch.float64, pin_memory=True)
xh.copy_(wl["x"]); wh.copy_(wl["w"]); torch.cuda.synchronize()
data = synth.ExpressionMatrix(xh.numpy(), wh.numpy(), wl["genes"], wl["cells"])
def once(sync):
    E.clear_device_cache(); torch.cuda.synchronize()
    t = [time.perf_counter()]
    def mark():
        if sync: torch.cuda.synchronize()
        t.append(time.perf_counter())
    real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5); mark()
    rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5); mark()
    res = S.sigs_vs_projections(wl["projections"], real, rnd, 0.33); mark()
    hs = next(iter(real.values()))._batch.host_scores(); mark()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])], 1e3 * (t[-1] - t[0])
once(True)
for sync in (True, False):
    parts, total = once(sync)
    print("sync=%d: score real %.1f | score rnd %.1f | sigs_vs_projections %.1f | D2H scores %.1f | tail %.1f | total %.1f ms"
          % ((sync,) + tuple(parts) + (total,)))
pr = cProfile.Profile(); pr.enable(); once(False); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
