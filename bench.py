#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the FastProject scoring core on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input of the
shape BASELINE.json's metric is quoted on (configs[1]: 20k cells x 15k genes,
500 signatures + 3000 random signatures, 6 projections):

    score all signatures  ->  rank  ->  per projection: fused Gaussian-
    neighbourhood contraction, median, background p-values.

Metric: signature x projection consistency pairs / s = (K+R)*P / t_step (whole
job).  ``value`` is timed with CUDA events with every input resident in HBM;
``e2e`` times the public Python mirror of the reference API with HOST (pinned)
inputs, host<->device copies inside the timed region.  The cell x signature
scores / s half of the metric is reported beside it (``scores_per_sec``).

N > 1 (torchrun): the signature blocks (real + random) are sharded over the
ranks, every rank holds the full expression matrix and all projections, and
the only exchange is an all-gather of the per-signature medians (weak scaling:
per-GPU work is fixed, the job has N x the signatures).

``--impl reference`` times the CPU restatement of the reference (the oracle
port; the reference is pure Python/NumPy and cannot travel to the GPU box) on a
bounded row-sample of the same workload, on the host cores of the box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: genes, cells, real sigs, random sigs per size (x6 sizes), projections
    "C2": dict(G=15000, N=20000, K=500, R_PER_SIZE=500, P=6),
    "C3": dict(G=20000, N=100000, K=1000, R_PER_SIZE=500, P=8),
    "C4": dict(G=2000, N=1000000, K=200, R_PER_SIZE=500, P=4),
    "tiny": dict(G=600, N=1024, K=24, R_PER_SIZE=10, P=2),
}
SIZES = [5, 10, 20, 50, 100, 200]
SIGMA = 0.33


def describe_config(workload, cfg, n_real, n_rnd, precision, world):
    """The ``config`` object of the JSON line -- identical for both arms."""
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    return {"workload": "%s: %d cells x %d genes, %d real + %d random signatures per GPU, "
                        "%d projections, sigma=0.33" % (workload, N, G, n_real, n_rnd, P),
            "precision_mode": precision, "signatures_per_gpu": n_real + n_rnd,
            "l2_policy": "inputs larger than L2 (X+W = %.1f GB per step)" % (G * N * 16 / 1e9),
            "parallelism": "signature-block x%d" % world}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FP_BENCH_WORKLOAD", "C2"))
    ap.add_argument("--precision", default=None, help="fp64 | tc (default: library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------
def make_workload(cfg, device, seed=1335607828, sig_seed=None):
    """Synthetic inputs of SURVEY.md 8d's recipe, generated on the device for
    speed (log1p-Gamma counts with ~40 % zeros, per-gene z-score, detection
    weights), signatures on the host with the reference's RNG calls."""
    import numpy as np
    import torch
    from fastproject_b200 import Signatures as S, synth
    G, N = cfg["G"], cfg["N"]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed % (2 ** 31))
    x = torch.empty((G, N), dtype=torch.float64, device=device)
    w = torch.empty((G, N), dtype=torch.float64, device=device)
    chunk = 1024
    for lo in range(0, G, chunk):
        hi = min(G, lo + chunk)
        shape = (hi - lo, N)
        # Gamma(shape .5, scale 4) = 4 * 0.5 * chi2_1 = 2 * z^2
        raw = torch.log1p(2.0 * torch.randn(shape, generator=gen, device=device, dtype=torch.float64) ** 2)
        raw = raw * (torch.rand(shape, generator=gen, device=device) < 0.6)
        u = torch.rand(shape, generator=gen, device=device, dtype=torch.float64)
        w[lo:hi] = torch.where(raw > 0, torch.ones_like(raw), u)
        mu = raw.mean(dim=1, keepdim=True)
        sd = raw.std(dim=1, keepdim=True, unbiased=False)
        sd[sd == 0] = 1.0
        x[lo:hi] = (raw - mu) / sd
    genes = ["G%05d" % i for i in range(G)]
    cells = ["C%06d" % i for i in range(N)]
    projections, _ = synth.make_projections(N, cfg["P"], seed=seed + 1)
    # one data set (expression, weights, projections) for the whole job; every rank owns its
    # own block of real + random signatures
    sig_seed = seed if sig_seed is None else sig_seed
    defs = synth.make_signature_dicts(genes, cfg["K"], seed=sig_seed + 2)
    real = [S.Signature(d, sg, "synthetic", nm) for nm, d, sg in defs]
    import random
    random.seed(sig_seed)
    np.random.seed(sig_seed % (2 ** 32))
    rnd = S.generate_random_sigs(genes, False, SIZES, cfg["R_PER_SIZE"])
    return dict(x=x, w=w, genes=genes, cells=cells, projections=projections, real=real, rnd=rnd)


# ----------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=smax,
                    power_w_max=(max(power) if power else None), samples=len(sm),
                    reasons=sorted(reasons))


# ----------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tflops=d.get("bf16_tflops_sustained", 1400.0),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at VERSION/INFO level; stdout carries the JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("FP_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from fastproject_b200 import Signatures as S, _engine as E, _native, _scoring, synth, distributed as D

    cfg = WORKLOADS[args.workload]
    precision = E.resolve_precision(args.precision, cfg["N"])
    # weak scaling: every rank owns its own block of K real + R random signatures
    wl = make_workload(cfg, device, seed=1335607828, sig_seed=1335607828 + 7919 * rank)
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    x, w = wl["x"], wl["w"]
    sigs = wl["real"] + wl["rnd"]
    kept, sig_ptr, sig_gene, sig_sign, counts = _scoring.pack_signatures(sigs, wl["genes"], 5)
    n_real = sum(1 for k in kept if k < len(wl["real"]))
    n_rows = len(kept)
    nnz = int(sig_ptr[-1])
    d_ptr = E.to_device(sig_ptr, torch.int32)
    d_gene = E.to_device(sig_gene, torch.int32)
    d_sign = E.to_device(sig_sign, torch.int8)
    proj_names = sorted(wl["projections"].keys())
    coords = [E.to_device(wl["projections"][p]) for p in proj_names]
    # who owns what: [n_real, n_rows] of every rank, and every row's numGenes
    shape_all = D.all_gather_counts([n_real, n_rows], device=device).reshape(world, 2)
    counts_all = D.all_gather_counts(counts, device=device)
    bg = D.GlobalBackground(counts_all, [int(v) for v in shape_all[:, 0]],
                            [int(v) for v in shape_all[:, 1]], device)
    ld = E.padded(N)
    dissim = torch.empty((n_rows, ld), dtype=torch.float64, device=device)
    sig2 = SIGMA ** 2

    ev = lambda: torch.cuda.Event(enable_timing=True)
    score_ev, cons_ev = [], []

    def step(record):
        e0, e1 = ev(), ev()
        e0.record()
        scores = E.score_signatures("weighted_avg", x, w, None, d_ptr, d_gene, d_sign, n_rows)
        e1.record()
        if record:
            score_ev.append((e0, e1))
        ranks = E.rank_rows(scores, N)
        meds = []
        for ic, c in enumerate(coords):
            c0, c1 = ev(), ev()
            c0.record()
            E.consistency(c, sig2, ranks, N, out=dissim, precision=precision, reuse_packed=ic > 0,
                          slot="bench")
            c1.record()
            if record:
                cons_ev.append((c0, c1))
            meds.append(E.row_medians(dissim, N))
        med = torch.stack(meds)                               # [P x rows]
        all_med = D.all_gather_rows(med)                      # the only exchange: P x rows doubles
        return bg.pvalues(all_med), all_med

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    t0, t1 = ev(), ev()
    sync()
    t0.record()
    for _ in range(args.steps):
        out = step(True)
    t1.record()
    sync()
    launches = _native.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t0.elapsed_time(t1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    score_ms = sum(a.elapsed_time(b) for a, b in score_ev) / max(1, len(score_ev))
    cons_ms = sum(a.elapsed_time(b) for a, b in cons_ev) / max(1, len(cons_ev))

    pairs = int(shape_all[:, 1].sum()) * P
    value = pairs / (ms_step * 1e-3)
    peaks = measured_peaks()
    cons_flops = 2.0 * N * N * (n_rows + 1)                   # per launch (one projection)
    score_bytes = G * N * 16.0 + N * n_rows * 8.0 + 5.0 * nnz  # per launch
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of
    # this exact configuration (profiles/r01_summary_slab.md); None where no capture exists
    # (tc: weight_planes_kernel 0.0043 + 1.5572 GB and consistency_mma_kernel 12.4369 + 0.6883 GB)
    ncu_traffic = {("C2", "tc", 1): 4314000 + 1557239000 + 12436882000 + 688290000,
                   ("C2", "fp64", 1): 603465984 + 551192064}
    roofline = dict(bound="tensor", kernel="fp_consistency[%s]" % precision,
                    achieved=cons_flops / (cons_ms * 1e-3) / 1e12, peak=peaks["tflops"],
                    unit="TFLOP/s", traffic=ncu_traffic.get((args.workload, precision, 1)),
                    peak_source=peaks["source"],
                    launch_ms=cons_ms, share_of_step=cons_ms * P / ms_step)
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    roofline_score = dict(bound="hbm", kernel="fp_score_signatures", unit="GB/s",
                          achieved=score_bytes / (score_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"],
                          launch_ms=score_ms, traffic=None,
                          gather_bytes=16.0 * nnz * N, share_of_step=score_ms / ms_step)
    roofline_score["frac"] = roofline_score["achieved"] / roofline_score["peak"]

    result = {
        "metric": "signature x projection consistency pairs/sec (full hot path: score + rank + "
                  "consistency + median + p-values)",
        "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if precision == "fp64" else "f64 + u8*s8->s32 digit planes (tcgen05 kind::i8)",
        "data": "synthetic",
        "config": describe_config(args.workload, cfg, n_real, n_rows - n_real, precision, world),
        "scores_per_sec": N * n_rows * world / (score_ms * 1e-3),
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_scoring": roofline_score,
        "clocks": clocks,
    }
    result["checksum_p"] = float(out[0].sum())
    result["checksum_median"] = float(out[1].sum())

    if not args.no_e2e:
        try:
            e2e = run_e2e(args, cfg, wl, precision, world, device)
        except torch.OutOfMemoryError as exc:          # e.g. C4: the device leg's buffers + two data sets
            if world > 1:
                raise
            e2e = {"unavailable": "out of device memory in the end-to-end leg: %s" % str(exc).split(".")[0]}
        if rank == 0:
            result["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(cfg, wl, budget_s=25.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(result)


def run_e2e(args, cfg, wl, precision, world=1, device=None):
    """Public API with HOST buffers: H2D of X and W from pinned memory, CSR
    packing, all kernels, D2H of the real signatures' scores and of the
    result matrices, every step.  With N ranks every rank runs the same calls on
    its own block of signatures (``distributed=True``: medians all-gathered inside
    ``sigs_vs_projections``); the time is the max over ranks."""
    import io
    import contextlib
    import torch
    import torch.distributed as dist
    from fastproject_b200 import Signatures as S, _engine as E, synth
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    import fastproject_b200 as FP
    # two sets of pinned host buffers: consecutive steps are different host arrays, so every
    # step uploads its own inputs; the upload of step k+1 is started at the beginning of step k
    # (fastproject_b200.prefetch) and runs behind the whole step
    host = []
    for _ in range(2):
        xh = torch.empty((G, N), dtype=torch.float64, pin_memory=True)
        wh = torch.empty((G, N), dtype=torch.float64, pin_memory=True)
        xh.copy_(wl["x"]); wh.copy_(wl["w"])
        host.append(synth.ExpressionMatrix(xh.numpy(), wh.numpy(), wl["genes"], wl["cells"]))
    torch.cuda.synchronize()
    # a pipeline needs a few steps to show its steady state; the first timed step still uploads
    # its matrices synchronously (its share shrinks with the number of steps)
    steps = max(args.steps, 12)
    warm = max(args.warmup, 3)
    d2h = 0

    def one_step(data, nxt):
        # this step's matrices first (a no-op when the previous step already started their
        # upload), then the NEXT step's: queued on the copy engine right behind, so the engine
        # never idles and never serves the later data set first
        FP.prefetch(data.base, data.weights, replicated=world > 1)
        if nxt is not None:
            FP.prefetch(nxt.base, nxt.weights, replicated=world > 1)
        real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5, distributed=world > 1)
        rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5, distributed=world > 1)
        res = S.sigs_vs_projections(wl["projections"], real, rnd, SIGMA, precision=precision,
                                    distributed=world > 1)
        first = next(iter(real.values()))
        host_scores = first._batch.host_scores()                 # D2H of all real scores
        E.expression_cache.evict(data.base)
        E.expression_cache.evict(data.weights)
        return real, rnd, host_scores.nbytes + res[2].nbytes + res[3].nbytes + res[4].nbytes

    # warm-up steps (untimed, pipelined like the timed ones: they also grow the pinned staging
    # pools and the allocator caches to their steady size), then everything resident is dropped
    # and ``steps`` timed steps run back to back; every timed step's inputs cross PCIe inside the
    # timed region (the first one synchronously, the others prefetched during the previous step)
    E.clear_device_cache()
    for it in range(warm):
        real, rnd, d2h = one_step(host[it % len(host)], host[(it + 1) % len(host)] if it + 1 < warm else None)
    E.clear_device_cache()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for it in range(steps):
        data = host[it % len(host)]
        nxt = host[(it + 1) % len(host)] if it + 1 < steps else None
        real, rnd, d2h = one_step(data, nxt)
        if os.environ.get("FP_BENCH_DEBUG"):
            sys.stderr.write("e2e step %d done at %.1f ms\n" % (it, 1e3 * (time.perf_counter() - t0)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax[0])
    times = [dt / steps] * steps
    E.clear_device_cache()
    t = sum(times) / len(times)
    n_rows = len(real) + len(rnd)
    nnz = sum(len(s.sig_dict) for s in wl["real"]) + sum(len(s.sig_dict) for s in wl["rnd"])
    # X and W are one data set: every rank uploads 1/world of their rows (all-gathered over NVLink)
    h2d = 2 * G * N * 8 + world * (nnz * 5 + (n_rows + 1) * 4 + P * 2 * N * 8)
    return {"value": world * n_rows * P / t, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h) * world, "s_per_step": t, "steps": len(times),
            "api": "Signatures.calculate_sig_scores x2 + Signatures.sigs_vs_projections(distributed=%s), "
                   "pinned host inputs; steps back to back, the upload of the next step's matrices "
                   "(fastproject_b200.prefetch at the start of the step) overlaps the current step%s"
                   % (world > 1, "; the expression matrices are one data set per job: every rank uploads "
                      "1/world of the rows over its PCIe link and the rows are all-gathered over NVLink"
                      if world > 1 else "")}


# ----------------------------------------------------------------------------
# CPU arm (oracle port of the reference)
# ----------------------------------------------------------------------------
def cpu_baseline(cfg, wl, budget_s=25.0, host=None):
    """Times the NumPy/SciPy restatement of the reference (oracle/fp_oracle.py)
    on a bounded sample and scales linearly to the full workload:
      scoring : the first S signatures on the full G x N matrices (cost is
                linear in the number of signatures);
      ranks   : scipy rankdata on S2 columns;
      consistency : ONE projection, a block of B rows of the N x N kernel
                against ALL K+R rank columns, with the row-blocked restatement
                (rows are independent and equal in cost -> x N/B x P);
      median  : np.median on the block's dissimilarities, scaled the same way.
    """
    import numpy as np
    from oracle import fp_oracle as O
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    if host is None:
        xh = wl["x"].cpu().numpy()
        wh = wl["w"].cpu().numpy()
    else:
        xh, wh = host
    odata = O.ExpressionMatrix(xh, wh, wl["genes"], wl["cells"])
    to_o = lambda s: O.Signature(s.sig_dict, s.signed, s.source, s.name)
    sample_sigs = [to_o(s) for s in (wl["real"][:24] + wl["rnd"][::max(1, len(wl["rnd"]) // 72)][:72])]
    t0 = time.perf_counter()
    res = O.calculate_sig_scores(odata, sample_sigs, "weighted_avg", None, 5)
    t_score = (time.perf_counter() - t0) / max(1, len(res))
    n_rows = len(wl["real"]) + len(wl["rnd"])
    t0 = time.perf_counter()
    rk = np.array([res[k].ranks for k in res])
    t_rank = (time.perf_counter() - t0) / max(1, len(res))
    # rank matrix of the full width: tile the sample's ranks (values do not change the cost)
    reps = (n_rows + rk.shape[0] - 1) // rk.shape[0]
    ranks = np.ascontiguousarray(np.tile(rk, (reps, 1))[:n_rows].T)        # N x (K+R)
    coords = wl["projections"][sorted(wl["projections"].keys())[0]]
    B = 512 if N >= 4096 else N
    # calibrate block size to the budget
    t0 = time.perf_counter()
    wt = O.neighbourhood_weights(coords, SIGMA, slice(0, B))
    dis = np.abs(ranks[:B] - np.dot(wt, ranks))
    t_blk = time.perf_counter() - t0
    rows_done = B
    while rows_done < N and (time.perf_counter() - t0) < budget_s * 0.6:
        lo = rows_done
        hi = min(N, lo + B)
        wt = O.neighbourhood_weights(coords, SIGMA, slice(lo, hi))
        dis = np.abs(ranks[lo:hi] - np.dot(wt, ranks))
        rows_done = hi
    t_cons_rows = (time.perf_counter() - t0) / rows_done                   # per row, one projection
    t0 = time.perf_counter()
    cols = min(n_rows, 64)
    full_col = np.abs(np.random.RandomState(0).normal(size=(N, cols)))
    np.median(full_col, axis=0)
    t_med = (time.perf_counter() - t0) / cols
    t_full = n_rows * (t_score + t_rank) + P * (N * t_cons_rows + n_rows * t_med)
    return {"value": n_rows * P / t_full, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
            "scores_per_sec": N / t_score,
            "sample": "oracle port (NumPy/SciPy float64, BLAS threads = all %d cores): scoring on %d "
                      "signatures; rankdata on the same; consistency on %d of %d rows of 1 of %d "
                      "projections against all %d signature columns (row-blocked restatement); "
                      "np.median on %d columns; scaled linearly to the full workload "
                      "(%.1f s extrapolated per step)" % (os.cpu_count() or 0, len(res), rows_done, N,
                                                          P, n_rows, cols, t_full),
            "extrapolated_s_per_step": t_full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from fastproject_b200 import Signatures as S, synth
    cfg = WORKLOADS[args.workload]
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    x, w, z, genes, cells = synth.make_expression(G, N)
    projections, _ = synth.make_projections(N, P, seed=1335607829)
    defs = synth.make_signature_dicts(genes, cfg["K"], seed=1335607830)
    import random
    random.seed(1335607828); np.random.seed(1335607828 % 2 ** 32)
    wl = dict(genes=genes, cells=cells, projections=projections,
              real=[S.Signature(d, sg, "synthetic", nm) for nm, d, sg in defs],
              rnd=S.generate_random_sigs(genes, False, SIZES, cfg["R_PER_SIZE"]))
    vals = []
    last = None
    for it in range(args.warmup + args.steps):
        budget = 20.0 if it >= args.warmup else 5.0
        last = cpu_baseline(cfg, wl, budget_s=budget, host=(x, w))
        if it >= args.warmup:
            vals.append(last["value"])
    value = sum(vals) / len(vals)
    n_rows = len(wl["real"]) + len(wl["rnd"])
    out = {"impl": "reference",
           "metric": "signature x projection consistency pairs/sec (full hot path: score + rank + "
                     "consistency + median + p-values)",
           "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * n_rows * P / value, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": describe_config(args.workload, cfg, len(wl["real"]), len(wl["rnd"]),
                                     args.precision or ("tc" if N <= 4000000 else "fp64"), args.gpus),
           "cpu_baseline": dict(last, value=value),
           "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_JSON_FD = None


def emit(obj):
    """The ONE JSON line of the contract goes to the real stdout; everything else that writes
    to fd 1 (NCCL's version banner, progress bars of the mirrored API) was sent to stderr."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
