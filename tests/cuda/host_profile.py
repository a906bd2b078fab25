"""Where the HOST time of calculate_sig_scores / sigs_vs_projections goes at the C2 shape
(cProfile; the GPU work is asynchronous, so what shows is Python + launch overhead)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import __graft_entry__ as ge

ge.build()
import bench
from fastproject_b200 import Signatures as S, synth

cfg = bench.WORKLOADS["C2"]
wl = bench.make_workload(cfg, torch.device("cuda"), seed=0)
data = synth.ExpressionMatrix(wl["x"], wl["w"], wl["genes"], wl["cells"])
for _ in range(2):
    real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5)
    rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5)
    S.sigs_vs_projections(wl["projections"], real, rnd, bench.SIGMA, precision="tc")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5)
    torch.cuda.synchronize()
pr.disable()
print("=== calculate_sig_scores(3000 random signatures) x 5")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    S.sigs_vs_projections(wl["projections"], real, rnd, bench.SIGMA, precision="tc")
pr.disable()
print("=== sigs_vs_projections x 3")
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
