"""Times fp_rank_columns on the BASELINE C2 shape (3 500 x 20 000) and on a 64 x 1 000 000 slab.
Run on the GPU box: python tests/cuda/rank_profile.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastproject_b200 import _engine as E  # noqa: E402


def time_it(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


for k, n, kind in ((3500, 20000, "normal"), (3500, 20000, "rounded"), (64, 1000000, "normal"), (15000, 3500, "zeros")):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
    if kind == "rounded":
        x = torch.round(x * 100) / 100
    if kind == "zeros":
        x = torch.where(x < 0.3, torch.zeros_like(x), x)
    ms = time_it(lambda: E.rank_rows(x, n))
    print("rank %6d x %8d %-8s %8.3f ms  (%.1f GB/s of scores+ranks)" % (k, n, kind, ms, 16.0 * k * n / ms / 1e6))
