"""Host profile of the hot path at BASELINE config 1's shape (2 000 cells x 5 000 genes, 100
signatures + the 3 000 background signatures, 8 projections): where sigs_vs_projections spends
its time when the kernels are small."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import __graft_entry__ as ge

ge.build()
import bench
from fastproject_b200 import Signatures as S, synth

cfg = dict(G=5000, N=2000, K=100, R_PER_SIZE=500, P=8)
wl = bench.make_workload(cfg, torch.device("cuda"), seed=0)
x, w = wl["x"].cpu().numpy(), wl["w"].cpu().numpy()
data = synth.ExpressionMatrix(x, w, wl["genes"], wl["cells"])
for rep in range(3):
    t0 = time.perf_counter()
    real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5)
    rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5)
    t1 = time.perf_counter()
    out = S.sigs_vs_projections(wl["projections"], real, rnd)
    t2 = time.perf_counter()
    print("pass %d: calculate_sig_scores x2 %.3f s, sigs_vs_projections %.3f s" % (rep, t1 - t0, t2 - t1))
pr = cProfile.Profile()
pr.enable()
real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5)
rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5)
out = S.sigs_vs_projections(wl["projections"], real, rnd)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
