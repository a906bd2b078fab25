#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the FastProject scoring core on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C3|C4]

A *step* is one pass of the hot path over ONE job of synthetic input:

    score all signatures  ->  rank  ->  per projection: Gaussian-neighbourhood contraction,
    median over the cells, background p-values.

The headline workload is BASELINE.json configs[1] (C2: 20k cells x 15k genes, 500 signatures +
3000 random signatures, 6 projections); the line also carries a ``"c3"`` block -- the north-star
shape (100k cells x 20k genes, 1000 + 3000 signatures, 8 projections) run the same way with
fewer steps.

Metric: signature x projection consistency pairs / s = (K + R) * P / t_step for the whole job.
``value`` is timed with CUDA events with every input resident in HBM; ``e2e`` times the public
Python mirror of the reference API with HOST (pinned) inputs, host<->device copies inside the
timed region.  The cell x signature scores / s half of the metric is ``scores_per_sec``.

N > 1 (torchrun, one process per GPU): the SAME job is split over the ranks ("scaling":
"strong") -- scoring and ranking by signature rows, the contraction by runs of the (projection
x signature-block) grid, with one all-gather of the rank rows and one all-reduce of the medians
(fastproject_b200/distributed.py).

CPU legs.  ``cpu_baseline`` (rank 0, N = 1): ONE WHOLE job of the same inputs on the host cores
-- the unmodified reference from baseline/_ref (kind "reference") when it is installed, else
the restatement in oracle/fp_oracle.py (kind "port") -- timed as a whole, not scaled from a
sample; its outputs are the checker of ``parity``.  ``--impl reference`` deals the units of one
whole job to the K timed steps (a step is 1/K of the job; the steps together are the job).
"""
from __future__ import annotations

import argparse
import hashlib
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
SEED = 1335607828
METRIC = ("signature x projection consistency pairs/sec (full hot path: score + rank + "
          "consistency + median + p-values)")


def describe_config(workload, cfg, n_real, n_rnd, precision, world):
    """The ``config`` object of the JSON line -- identical for both arms."""
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    return {"workload": "%s: %d cells x %d genes, %d real + %d random signatures, %d projections, "
                        "sigma=0.33; one job for all GPUs" % (workload, N, G, n_real, n_rnd, P),
            "precision_mode": precision, "signatures": n_real + n_rnd,
            "l2_policy": "inputs larger than L2 (X+W = %.1f GB per step)" % (G * N * 16 / 1e9),
            "parallelism": "1 GPU" if world == 1 else
                           "%d GPUs: signature rows for scoring/ranking, (projection x signature-block) "
                           "runs for the contraction" % world}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FP_BENCH_WORKLOAD", "C2"))
    ap.add_argument("--precision", default=None, help="fp64 | tc (default: library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 block")
    ap.add_argument("--c3-steps", type=int, default=3)
    ap.add_argument("--no-c5", action="store_true", help="skip the scoring-only sweep")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------
def make_signatures(cfg, genes, seed=SEED):
    import random
    import numpy as np
    from fastproject_b200 import Signatures as S, synth
    defs = synth.make_signature_dicts(genes, cfg["K"], seed=seed + 2)
    real = [S.Signature(d, sg, "synthetic", nm) for nm, d, sg in defs]
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    rnd = S.generate_random_sigs(genes, False, SIZES, cfg["R_PER_SIZE"])
    return real, rnd


def make_workload(cfg, device, seed=SEED, keep_raw=False):
    """Synthetic inputs of SURVEY.md 8d's recipe, generated on the device for speed: raw
    log1p-Gamma counts with ~40 % zeros; per-cell logistic false-negative curves; then the two
    matrices the path reads, made the way the pipeline makes them (Pipelines.py:114-120, 202):
    W = compute_weights(raw, curves), X = per-gene z-score of raw.  Signatures on the host with
    the reference's RNG calls.  Every rank builds the same job."""
    import numpy as np
    import torch
    from fastproject_b200 import _engine as E, synth
    G, N = cfg["G"], cfg["N"]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed % (2 ** 31))
    projections, gradient = synth.make_projections(N, cfg["P"], seed=seed + 1)
    # a gradient along the first projection planted into the first 150 genes (the first five
    # real signatures draw from them): some signature x projection pairs are genuinely
    # consistent, so the significance calls of the parity check are not all "no"
    planted = torch.from_numpy(np.maximum(gradient, 0.0)).to(device)
    raw = torch.empty((G, N), dtype=torch.float64, device=device)
    chunk = 1024
    for lo in range(0, G, chunk):
        hi = min(G, lo + chunk)
        shape = (hi - lo, N)
        # Gamma(shape .5, scale 4) = 4 * 0.5 * chi2_1 = 2 * z^2
        r = torch.log1p(2.0 * torch.randn(shape, generator=gen, device=device, dtype=torch.float64) ** 2)
        if lo < 150:
            k = min(hi, 150) - lo
            r[:k] += planted[None, :] * (0.5 + torch.rand((k, 1), generator=gen, device=device, dtype=torch.float64))
        raw[lo:hi] = r * (torch.rand(shape, generator=gen, device=device) < 0.6)
    rng = np.random.RandomState(seed % (2 ** 31))
    params = np.stack([rng.uniform(0.5, 2.5, N), rng.uniform(0.5, 3.0, N), np.zeros(N), np.ones(N)])
    w = E.compute_weights(raw, E.to_device(params))
    x = E.normalize_rows(raw)
    genes = ["G%05d" % i for i in range(G)]
    cells = ["C%06d" % i for i in range(N)]
    real, rnd = make_signatures(cfg, genes, seed)
    return dict(x=x, w=w, raw=raw if keep_raw else None, params=params, genes=genes, cells=cells,
                projections=projections, real=real, rnd=rnd)


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
# B200 arm: the job with every input resident in HBM
# ----------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tflops=d.get("bf16_tflops_sustained", 1400.0),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


def captured_traffic(workload, precision, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` from the committed
    ``ncu --set full`` capture of this workload (profiles/r02_traffic.json names the capture file
    and its sha256); None when no capture of this configuration is committed."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as fh:
        table = json.load(fh)
    entry = table.get("%s/%s/%s" % (workload, precision, kernel))
    if not entry:
        return None, None
    return entry.get("dram_bytes_per_launch"), entry.get("source")


class ResidentJob(object):
    """One job on this rank's share of the work; inputs resident in HBM."""

    def __init__(self, name, cfg, wl, device, world, rank, precision):
        import numpy as np
        import torch
        from fastproject_b200 import Signatures as S, _engine as E, _scoring, distributed as D
        self.name, self.cfg, self.wl, self.device, self.world, self.rank = name, cfg, wl, device, world, rank
        self.E, self.D, self.S, self.torch = E, D, S, torch
        self.precision = precision
        G, N, P = cfg["G"], cfg["N"], cfg["P"]
        sigs = wl["real"] + wl["rnd"]
        kept, sig_ptr, sig_gene, sig_sign, counts = _scoring.pack_signatures(sigs, wl["genes"], 5)
        self.n_real = sum(1 for k in kept if k < len(wl["real"]))
        self.n_rows = len(kept)
        self.counts = counts
        self.nnz = int(sig_ptr[-1])
        # stage 1: my rows of the signature batch, dealt round-robin (same mix of sizes on every rank)
        self.owner_rows = D.cyclic_owner_rows(self.n_rows, world)
        mine = self.owner_rows[rank]
        self.rows = mine
        starts, stops = sig_ptr[mine].astype(np.int64), sig_ptr[mine + 1].astype(np.int64)
        lens = stops - starts
        local_ptr = np.zeros(mine.size + 1, dtype=np.int64)
        np.cumsum(lens, out=local_ptr[1:])
        take = np.concatenate([np.arange(a, b) for a, b in zip(starts, stops)]) if mine.size else np.zeros(0, np.int64)
        self.nnz_local = int(local_ptr[-1])
        self.d_ptr = E.to_device(local_ptr.astype(np.int32), torch.int32)
        self.d_gene = E.to_device(sig_gene[take], torch.int32)
        self.d_sign = E.to_device(sig_sign[take], torch.int8)
        # stage 2: my runs of the (projection x signature-block) grid
        self.proj_names = sorted(wl["projections"].keys())
        self.coords = E.to_device(np.stack([wl["projections"][p] for p in self.proj_names]))
        if world > 1:
            self.runs = D.partition_units(P, self.n_rows, world)[rank]
        else:
            self.runs = [(p, 0, self.n_rows) for p in range(P)]
        order, gptr, grp = S._background_groups(list(counts[self.n_real:]), list(counts[:self.n_real]))
        self.d_order = E.to_device(order, torch.int32)
        self.d_gptr = E.to_device(gptr, torch.int32)
        self.d_grp = E.to_device(grp, torch.int32)
        self.ld = E.padded(N)
        widest = max([h - l for _, l, h in self.runs] or [1])
        self.overlap = E.MedianOverlap(widest, self.ld, device)
        self.dissim = self.overlap.bufs[0]
        self.med_all = torch.zeros((P, self.n_rows), dtype=torch.float64, device=device)
        self.p_all = torch.empty((P, self.n_real), dtype=torch.float64, device=device)
        self.sig2 = SIGMA ** 2
        self.events = []              # (stage, start, stop) of the recorded steps

    def _ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def step(self, record=False):
        E, D, torch = self.E, self.D, self.torch
        N, P = self.cfg["N"], self.cfg["P"]
        marks = []

        def mark(stage):
            if record:
                e = self._ev()
                e.record()
                marks.append((stage, e))

        self.last_scores = self.last_ranks = None       # the previous step's outputs go before new ones come
        mark("begin")
        scores = E.score_signatures("weighted_avg", self.wl["x"], self.wl["w"], None, self.d_ptr, self.d_gene,
                                    self.d_sign, int(self.rows.size))
        mark("score")
        ranks = E.rank_rows(scores, N)
        mark("rank")
        if self.world > 1:
            ranks = D.all_gather_rank_rows(ranks, self.owner_rows, N)
            self.med_all.zero_()
        mark("gather")
        # (what Signatures._contract_runs does) row ranges met by several projections are packed once
        uses, packed = dict(), dict()
        for _, lo, hi in self.runs:
            uses[(lo, hi)] = uses.get((lo, hi), 0) + 1
        for p, lo, hi in self.runs:
            if uses[(lo, hi)] > 1 and (lo, hi) not in packed:
                packed[(lo, hi)] = E.pack_values(ranks[lo:hi], N, precision=self.precision)
            # the median of this run goes to a side stream and overlaps the next run's weight planes
            dis = E.consistency(self.coords[p], self.sig2, ranks[lo:hi], N, out=self.overlap.buffer(hi - lo),
                                precision=self.precision, slot="bench", any_cell_order=True,
                                packed=packed.get((lo, hi)))
            mark("consistency")
            self.overlap.median_into(dis, N, self.med_all[p, lo:hi])
        self.overlap.join()
        mark("median (last one; the others overlap the next contraction)")
        if self.world > 1:
            D.all_reduce_sum(self.med_all)
        for i in range(P):
            p, _, _ = E.background_pvalues(self.med_all[i, :self.n_real].contiguous(), self.d_grp,
                                           self.med_all[i, self.n_real:].contiguous(), self.d_order, self.d_gptr)
            self.p_all[i] = p
        mark("reduce+pvalues")
        if record:
            self.events.append(marks)
        self.last_scores, self.last_ranks = scores, ranks
        return self.p_all, self.med_all

    def stage_ms(self):
        """Mean milliseconds per step of every stage, from the recorded events (this rank)."""
        out = dict()
        for marks in self.events:
            for (_, a), (stage, b) in zip(marks[:-1], marks[1:]):
                out[stage] = out.get(stage, 0.0) + a.elapsed_time(b)
        n = max(1, len(self.events))
        return dict((k, v / n) for k, v in out.items())

    def results(self):
        """Host copies in the reference's layout: consistency [K x P], log10 q [K x P],
        background consistency [R x P]."""
        import numpy as np
        N = self.cfg["N"]
        med = self.med_all.cpu().numpy()
        cons = 1 - med[:, :self.n_real].T / N
        rnd_cons = 1 - med[:, self.n_real:].T / N
        q = self.S.p_to_q(self.p_all.cpu().numpy().T.copy())
        q[q == 0] = 1e-300
        with np.errstate(invalid="ignore", divide="ignore"):
            logq = np.log10(q)
        return cons, logq, rnd_cons


def timed_run(job, steps, warmup, world, sampler=None):
    """W warm-up steps, then exactly K steps between barrier + synchronize; CUDA events, max over
    ranks.  Returns (ms per step, launches in the timed region, clocks)."""
    import torch
    import torch.distributed as dist
    from fastproject_b200 import _native

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        job.step(False)
    sync()
    if sampler is not None:
        sampler.start()
    launches0 = _native.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0.record()
    for _ in range(steps):
        job.step(True)
    t1.record()
    sync()
    launches = _native.launch_count() - launches0
    clocks = sampler.stop() if sampler is not None else None
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=job.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if job.runs:
        _, lo, hi = job.runs[-1]
        if job.E.consistency_refused(job.last_ranks[lo:hi], job.cfg["N"], precision=job.precision, slot="bench"):
            raise RuntimeError("the tensor-core path refused the rank values: the timed results are NaN")
    return float(t[0]) / steps, int(launches), clocks


def rooflines(job, ms_step, workload):
    """roofline of the dominant kernel (the contraction: tensor pipe) and of the scoring kernel
    (HBM), from this rank's stage times."""
    cfg = job.cfg
    G, N = cfg["G"], cfg["N"]
    peaks = measured_peaks()
    st = job.stage_ms()
    run_rows = sum(hi - lo + 1 for _, lo, hi in job.runs)            # + the ones row of every launch
    cons_ms = st.get("consistency", 0.0)                              # all of this rank's launches, per step
    n_launch = max(1, len(job.runs))
    cons_flops = 2.0 * N * N * run_rows
    traffic, traffic_src = captured_traffic(workload, job.precision, "fp_consistency")
    roof = dict(bound="tensor", kernel="fp_consistency[%s]" % job.precision,
                achieved=cons_flops / max(cons_ms, 1e-9) / 1e9, peak=peaks["tflops"], unit="TFLOP/s",
                traffic=traffic, traffic_source=traffic_src, peak_source=peaks["source"],
                launch_ms=cons_ms / n_launch, launches_per_step=n_launch,
                share_of_step=cons_ms / ms_step,
                algorithmic="2*N^2 flop per (signature, projection) pair; one launch = one projection x "
                            "(this rank's rows + 1 normaliser row)")
    roof["frac"] = roof["achieved"] / roof["peak"]
    rows_local = int(job.rows.size)
    score_ms = st.get("score", 0.0)
    score_bytes = G * N * 16.0 + N * rows_local * 8.0 + 5.0 * job.nnz_local
    s_traffic, s_src = captured_traffic(workload, job.precision, "fp_score_signatures")
    roof_s = dict(bound="hbm", kernel="fp_score_signatures", unit="GB/s",
                  achieved=score_bytes / max(score_ms, 1e-9) / 1e6, peak=peaks["hbm_gbs"],
                  launch_ms=score_ms, traffic=s_traffic, traffic_source=s_src,
                  gather_bytes=16.0 * job.nnz_local * N, share_of_step=score_ms / ms_step,
                  algorithmic="X and W once (16 B per gene x cell) + 8 B per score + 5 B per signature gene")
    roof_s["frac"] = roof_s["achieved"] / roof_s["peak"]
    return roof, roof_s, st


def compare_with_cpu(job, cpu):
    """``parity`` of the timed workload: the GPU job's outputs against the whole CPU job's."""
    import numpy as np
    cons, logq, rnd_cons = job.results()
    scores = job.E.to_host(job.last_scores[:, :job.cfg["N"]])            # N = 1: all rows
    ref_scores = np.concatenate([cpu.scores_matrix("real"), cpu.scores_matrix("rnd")], axis=0)
    out = {"checker": "whole job on the host: " + cpu.kind, "cells_checked": job.cfg["N"],
           "signatures_checked": int(ref_scores.shape[0]), "projections_checked": len(job.proj_names),
           "scores_bit_identical": bool(scores.shape == ref_scores.shape and np.array_equal(scores, ref_scores))}
    rel = np.abs(cons - cpu.cons) / np.abs(cpu.cons)
    rel_r = np.abs(rnd_cons - cpu.rnd_cons) / np.abs(cpu.rnd_cons)
    out["max_rel_err"] = float(max(rel.max(), rel_r.max()))
    out["tolerance"] = 1e-6
    both = np.isfinite(logq) & np.isfinite(cpu.logq)
    out["max_abs_err_log10q"] = float(np.abs(logq[both] - cpu.logq[both]).max()) if both.any() else None
    out["ranking_identical"] = bool(np.array_equal(np.argsort(logq, axis=None, kind="stable"),
                                                   np.argsort(cpu.logq, axis=None, kind="stable")))
    out["significance_calls_identical"] = bool(np.array_equal(logq < -1.3, cpu.logq < -1.3))
    out["significant_pairs"] = int((cpu.logq < -1.3).sum())
    out["ok"] = bool(out["scores_bit_identical"] and out["max_rel_err"] <= 1e-6 and out["ranking_identical"]
                     and out["significance_calls_identical"])
    return out


def subset_parity(job, n_sigs=16, n_cells=1024):
    """Parity on a SUBSET for shapes whose whole job the host cannot do in minutes (C3; and any
    N > 1 run): the first signatures of this rank's first run are scored and ranked by the oracle
    from the gene rows they use (bit-exact comparison), and ``n_cells`` rows of the N x N kernel
    of that projection are contracted by the row-blocked oracle and compared with the GPU's
    |rank - prediction| for those cells."""
    import numpy as np
    import torch
    from oracle import fp_oracle as O
    if not job.runs or job.rank != 0:
        return None
    p, lo, hi = job.runs[0]
    N = job.cfg["N"]
    # rank 0's first signatures (its rows are every world-th row of the job) inside its first run
    rows = job.rows[:n_sigs]
    rows = rows[(rows >= lo) & (rows < hi)]
    take = int(rows.size)
    sigs = (job.wl["real"] + job.wl["rnd"])
    from fastproject_b200 import _scoring
    kept, _, _, _, _ = _scoring.pack_signatures(sigs, job.wl["genes"], 5)
    chosen = [sigs[kept[r]] for r in rows]
    gene_pos = dict((g, i) for i, g in enumerate(job.wl["genes"]))
    grows = sorted(set(gene_pos[g] for s in chosen for g in s.sig_dict if g in gene_pos))
    idx = torch.tensor(grows, dtype=torch.long, device=job.device)
    xs = job.wl["x"].index_select(0, idx).cpu().numpy()
    ws = job.wl["w"].index_select(0, idx).cpu().numpy()
    odata = O.ExpressionMatrix(xs, ws, [job.wl["genes"][r] for r in grows], job.wl["cells"])
    ores = O.calculate_sig_scores(odata, [O.Signature(s.sig_dict, s.signed, s.source, s.name) for s in chosen],
                                  "weighted_avg", None, 5)
    want_scores = np.array([ores[s.name].scores for s in chosen])
    want_ranks = np.array([ores[s.name].ranks for s in chosen])
    # GPU: scores of those rows (the first of rank 0's local rows), their ranks in the gathered
    # matrix, and the contraction of the run they belong to
    rows_dev = torch.from_numpy(rows).to(job.device)
    got_scores = job.last_scores[:take, :N].cpu().numpy()
    got_ranks = job.last_ranks.index_select(0, rows_dev)[:, :N].cpu().numpy()
    dis = job.E.consistency(job.coords[p], job.sig2, job.last_ranks[lo:hi], N, out=job.dissim[:hi - lo],
                            precision=job.precision, slot="bench")
    cells = np.sort(np.random.RandomState(11).choice(N, size=min(n_cells, N), replace=False))
    got = dis.index_select(0, rows_dev - lo).index_select(1, torch.from_numpy(cells).to(job.device)).cpu().numpy()
    coords = job.wl["projections"][job.proj_names[p]]
    want = np.empty_like(got)
    vt = np.ascontiguousarray(want_ranks.T)
    for b in range(0, cells.size, 256):
        sel = cells[b:b + 256]
        wt = O.neighbourhood_weights(coords, SIGMA, sel)
        want[:, b:b + 256] = np.abs(vt[sel] - np.dot(wt, vt)).T
    err = float(np.abs(got - want).max())
    return {"checker": "oracle port on a subset (scores and ranks from the gene rows used; row-blocked "
                       "N x N kernel)", "signatures_checked": int(take), "cells_checked": int(cells.size),
            "projection": job.proj_names[p],
            "scores_bit_identical": bool(np.array_equal(got_scores, want_scores)),
            "ranks_bit_identical": bool(np.array_equal(got_ranks, want_ranks)),
            "max_abs_err_rank_units": err, "max_rel_err": err / N, "tolerance": 1e-6,
            "ok": bool(np.array_equal(got_scores, want_scores) and np.array_equal(got_ranks, want_ranks)
                       and err / N <= 1e-6)}


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
        # NCCL_DEBUG is the caller's: whatever NCCL prints goes to stderr (fd 1 is the JSON line's)
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from fastproject_b200 import _engine as E

    cfg = WORKLOADS[args.workload]
    precision = E.resolve_precision(args.precision, cfg["N"])
    want_e2e = not args.no_e2e
    wl = make_workload(cfg, device, keep_raw=want_e2e)
    job = ResidentJob(args.workload, cfg, wl, device, world, rank, precision)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step, launches, clocks = timed_run(job, args.steps, args.warmup, world, sampler)
    if os.environ.get("FP_BENCH_DEBUG"):
        sys.stderr.write("peak device memory after the timed run: %.1f GB\n" % (torch.cuda.max_memory_allocated() / 1e9))
    P = cfg["P"]
    roof, roof_s, stages = rooflines(job, ms_step, args.workload)
    result = {
        "metric": METRIC, "value": job.n_rows * P / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if precision == "fp64" else "f64 + u8*s8->s32 digit planes (tcgen05 kind::i8)",
        "data": "synthetic",
        "config": describe_config(args.workload, cfg, job.n_real, job.n_rows - job.n_real, precision, world),
        "scores_per_sec": cfg["N"] * job.n_rows / (max(stages.get("score", 0.0), 1e-9) * 1e-3),
        "gpu_launches": launches, "roofline": roof, "roofline_scoring": roof_s,
        "stage_ms_rank0": stages, "clocks": clocks,
    }
    out = job.step(False)
    result["checksum_p"] = float(out[0].sum())
    result["checksum_median"] = float(out[1].sum())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"], cpu = cpu_whole_job(cfg, wl)
        result["parity"] = compare_with_cpu(job, cpu)
        del cpu
    elif not args.no_cpu_baseline:
        par = subset_parity(job)
        if rank == 0:
            result["parity"] = par
    if world > 1:
        dist.barrier()

    if want_e2e:
        try:
            e2e = run_e2e(args, cfg, wl, precision, world, device)
        except torch.OutOfMemoryError as exc:          # e.g. C4: the device leg's buffers + two data sets
            if world > 1:
                raise
            e2e = {"unavailable": "out of device memory in the end-to-end leg: %s" % str(exc).split(".")[0]}
        if rank == 0:
            result["e2e"] = e2e

    # the north-star shape, same arm, fewer steps
    if not args.no_c3 and args.workload == "C2":
        del job
        for k in ("x", "w", "raw"):
            wl[k] = None
        E.clear_device_cache()
        torch.cuda.empty_cache()
        c3 = run_block("C3", args, device, world, rank, args.c3_steps, 1)
        if rank == 0:
            result["c3"] = c3
    # BASELINE configs[4]: signature-scoring-only sweep (cell x signature scores/s against the HBM roofline)
    if not args.no_c5 and args.workload == "C2" and world == 1:
        E.clear_device_cache()
        torch.cuda.empty_cache()
        result["c5"] = run_scoring_sweep(device)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(result)


C5_SHAPES = [  # genes, cells, signatures
    (15000, 10000, 1000), (15000, 10000, 10000), (15000, 100000, 1000), (15000, 100000, 10000),
    (15000, 100000, 50000), (2000, 1000000, 1000), (2000, 1000000, 10000),
]


def run_scoring_sweep(device):
    """fp_score_signatures alone (weighted_avg, random signatures of the reference's six sizes) on
    resident matrices: time per launch (CUDA events, best of 2 after a warm-up; the matrices are
    larger than L2), scores/s, and the fraction of the HBM roofline on the compulsory bytes
    16 G N + 8 N K + 5 nnz."""
    import gc
    import random
    import numpy as np
    import torch
    from fastproject_b200 import Signatures as S, _engine as E, _scoring
    peaks = measured_peaks()
    rows = []
    for G, N, K in C5_SHAPES:
        gen = torch.Generator(device=device)
        gen.manual_seed(G + N)
        x = torch.randn((G, N), dtype=torch.float64, device=device, generator=gen)
        w = torch.rand((G, N), dtype=torch.float64, device=device, generator=gen)
        genes = ["G%05d" % i for i in range(G)]
        random.seed(1)
        np.random.seed(1)
        sigs = S.generate_random_sigs(genes, False, SIZES, (K + 5) // 6)[:K]
        kept, sp, sg, ss, _ = _scoring.pack_signatures(sigs, genes, 5)
        dp, dg, ds = E.to_device(sp, torch.int32), E.to_device(sg, torch.int32), E.to_device(ss, torch.int8)
        E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for r in range(2):
            E.score_signatures("weighted_avg", x, w, None, dp, dg, ds, len(kept))
            ev[r + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(2))
        nnz = int(sp[-1])
        compulsory = G * N * 16.0 + N * len(kept) * 8.0 + 5.0 * nnz
        rows.append({"genes": G, "cells": N, "signatures": len(kept), "nnz": nnz, "ms": ms,
                     "scores_per_sec": N * len(kept) / (ms * 1e-3),
                     "compulsory_GBps": compulsory / ms / 1e6,
                     "frac_of_hbm_roofline": compulsory / ms / 1e6 / peaks["hbm_gbs"],
                     "gather_GBps": 16.0 * nnz * N / ms / 1e6})
        del x, w, dp, dg, ds
        gc.collect()
        torch.cuda.empty_cache()
    return {"kernel": "fp_score_signatures (weighted_avg, gather kernel)", "peak_GBps": peaks["hbm_gbs"],
            "note": "gather_GBps = 16 nnz N / t is the L2 -> SM rate of the gathers (cap of the part: ~12 TB/s)",
            "rows": rows}


def run_block(name, args, device, world, rank, steps, warmup):
    """Another workload through the same resident arm; returns a JSON-ready block."""
    import torch.distributed as dist
    from fastproject_b200 import _engine as E
    cfg = WORKLOADS[name]
    precision = E.resolve_precision(args.precision, cfg["N"])
    wl = make_workload(cfg, device)
    job = ResidentJob(name, cfg, wl, device, world, rank, precision)
    ms_step, launches, _ = timed_run(job, steps, warmup, world, None)
    roof, roof_s, stages = rooflines(job, ms_step, name)
    block = {"value": job.n_rows * cfg["P"] / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world,
             "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "scaling": "strong",
             "config": describe_config(name, cfg, job.n_real, job.n_rows - job.n_real, precision, world),
             "scores_per_sec": cfg["N"] * job.n_rows / (max(stages.get("score", 0.0), 1e-9) * 1e-3),
             "gpu_launches": launches, "roofline": roof, "roofline_scoring": roof_s, "stage_ms_rank0": stages}
    if not args.no_cpu_baseline:
        par = subset_parity(job)
        if rank == 0:
            block["parity"] = par
            if world == 1:
                block["cpu_baseline"] = cpu_sampled(cfg, wl, budget_s=20.0)
    if world > 1:
        dist.barrier()
    return block


def run_e2e(args, cfg, wl, precision, world=1, device=None):
    """Public API with HOST buffers, the pipeline's own order (Pipelines.py:114-120, 202-253):
    H2D of the raw expression matrix from pinned memory, false-negative weights and per-gene
    z-scores on the device (Transforms.compute_weights / NormalizationMethods.row_normalization
    mirrors), signature matching + CSR packing on the host, scoring, ranks, all projections,
    D2H of the real signatures' scores and of the result matrices -- every step.  With N ranks
    every rank makes the same calls (``distributed=True``): ONE job, the time is the max over
    ranks."""
    import torch
    import torch.distributed as dist
    import fastproject_b200 as FP
    from fastproject_b200 import Signatures as S, Transforms as T, NormalizationMethods as NM, _engine as E, synth
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    # two pinned host buffers: consecutive steps are different host arrays, so every step uploads
    # its own input; the upload of step k+1 is started at the beginning of step k
    from fastproject_b200 import distributed as D
    dist_mode = world > 1
    rank = dist.get_rank() if dist_mode else 0
    row_lo, row_hi = D.shard_bounds(G, world, rank)
    host = []
    for _ in range(2):
        # every rank holds (and uploads) only its block of gene rows of the job's matrix
        rh = torch.empty((row_hi - row_lo, N), dtype=torch.float64, pin_memory=True)
        rh.copy_(wl["raw"][row_lo:row_hi])
        host.append(rh.numpy())
    wl["raw"] = None
    params = wl["params"]
    torch.cuda.synchronize()
    steps = max(args.steps, 12)
    warm = max(args.warmup, 3)
    trace = [] if os.environ.get("FP_BENCH_DEBUG") else None

    def stamp(what):
        if trace is not None:
            trace.append((what, time.perf_counter()))

    if trace is not None:                      # who stalls the host: collector passes, pinned allocations
        import gc
        started = {}

        def on_gc(phase, info):
            if phase == "start":
                started["t"] = time.perf_counter()
            elif time.perf_counter() - started.get("t", 0) > 2e-3:
                sys.stderr.write("e2e gc gen%d %.1f ms (%d collected)\n"
                                 % (info["generation"], 1e3 * (time.perf_counter() - started["t"]), info["collected"]))
        gc.callbacks.append(on_gc)
        real_slot = E._pinned_slot

        def traced_slot(nbytes):
            n0, t0 = len(E._small_pool), time.perf_counter()
            slot = real_slot(nbytes)
            if len(E._small_pool) != n0:
                sys.stderr.write("e2e new pinned slot of %d bytes: %.1f ms (pool %d)\n"
                                 % (slot[0].numel(), 1e3 * (time.perf_counter() - t0), len(E._small_pool)))
            return slot
        E._pinned_slot = traced_slot

    def one_step(raw, nxt):
        stamp("begin")
        FP.prefetch(raw)
        if nxt is not None:
            FP.prefetch(nxt)
        raw_dev = E.expression_cache.get(raw)
        # weights and z-scores are per-gene-row operations: each rank makes them for its rows
        w_dev = T.compute_weights(None, params, raw_dev)                       # Pipelines.py:120
        x_dev = NM.row_normalization(raw_dev)                                  # Pipelines.py:202
        if dist_mode:                                                          # all rows to every GPU over NVLink
            bounds = D.even_bounds(G, world)
            w_dev = D.all_gather_row_blocks(w_dev, bounds)
            x_dev = D.all_gather_row_blocks(x_dev, bounds)
        data = synth.ExpressionMatrix(x_dev, w_dev, wl["genes"], wl["cells"])
        stamp("prep issued")
        real = S.calculate_sig_scores(data, wl["real"], "weighted_avg", None, 5, distributed=dist_mode)
        S.prefetch_scores(real)                  # their D2H runs beside the rest of the step
        stamp("score real issued")
        rnd = S.calculate_sig_scores(data, wl["rnd"], "weighted_avg", None, 5, distributed=dist_mode)
        stamp("score rnd issued")
        res = S.sigs_vs_projections(wl["projections"], real, rnd, SIGMA, precision=precision,
                                    distributed=dist_mode)
        stamp("svp returned")
        first = next(iter(real.values()))
        host_scores = first._batch.host_scores_local()           # D2H of this rank's real scores
        FP.evict(raw)
        stamp("scores on host")
        return real, rnd, host_scores.nbytes + res[2].nbytes + res[3].nbytes + res[4].nbytes

    E.clear_device_cache()
    for it in range(warm):
        real, rnd, d2h = one_step(host[it % 2], host[(it + 1) % 2] if it + 1 < warm else None)
    E.clear_device_cache()
    # the collector stays on, but what exists now (the interpreter's modules, the workload, the
    # CPU checker's job: about a million container objects, 30-170 ms per full pass) is moved out
    # of its generations, so a full pass inside the timed region walks only the steps' own objects
    import gc
    gc.collect()
    gc.freeze()
    torch.cuda.synchronize()
    if dist_mode:
        dist.barrier()
    t0 = time.perf_counter()
    for it in range(steps):
        real, rnd, d2h = one_step(host[it % 2], host[(it + 1) % 2] if it + 1 < steps else None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gc.unfreeze()
    if trace is not None:
        for (a, ta), (b, tb) in zip(trace[:-1], trace[1:]):
            if b != "begin":
                sys.stderr.write("e2e %-18s %7.2f ms\n" % (b, 1e3 * (tb - ta)))
            else:
                sys.stderr.write("e2e ---- step gap   %7.2f ms\n" % (1e3 * (tb - ta)))
    tsum = torch.tensor([dt, float(d2h)], dtype=torch.float64, device=device)
    if dist_mode:
        tmax = tsum.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, d2h = float(tmax[0]), float(tsum[1])
    E.clear_device_cache()
    t = dt / steps
    n_rows = len(real) + len(rnd)
    nnz = sum(len(s.sig_dict) for s in wl["real"]) + sum(len(s.sig_dict) for s in wl["rnd"])
    # the raw matrix is one data set: every rank uploads its 1/world of the gene rows
    h2d = G * N * 8 + nnz * 5 + (n_rows + world) * 4 + world * (P * 2 * N * 8 + 4 * N * 8)
    return {"value": n_rows * P / t, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "s_per_step": t, "steps": steps,
            "api": "fastproject_b200.prefetch(raw) -> Transforms.compute_weights + NormalizationMethods."
                   "row_normalization on the device -> Signatures.calculate_sig_scores x2 (+ prefetch_scores) -> Signatures."
                   "sigs_vs_projections(distributed=%s); pinned host input; steps back to back, the upload of "
                   "the next step's matrix overlaps the current step; gc.freeze() before the timed loop%s"
                   % (dist_mode, "; one job: every rank uploads its 1/world of the gene rows over its own PCIe "
                      "link, makes weights and z-scores for them, and X / W are all-gathered over NVLink"
                      if dist_mode else "")}


# ----------------------------------------------------------------------------
# CPU legs
# ----------------------------------------------------------------------------
def cpu_whole_job(cfg, wl, host=None):
    """ONE WHOLE job on the host cores: the installed reference (kind "reference") or the port.
    Returns (cpu_baseline dict, the finished CpuJob)."""
    from oracle import cpu_job as J
    P = cfg["P"]
    if host is None:
        xh, wh = wl["x"].cpu().numpy(), wl["w"].cpu().numpy()
    else:
        xh, wh = host
    job = J.CpuJob(xh, wh, wl["genes"], wl["cells"], wl["projections"], wl["real"], wl["rnd"], SIGMA, 5)
    with J.blas_threads() as threads:
        t0 = time.perf_counter()
        job.run()
        dt = time.perf_counter() - t0
    n_rows = len(job.real) + len(job.rnd)
    return ({"value": n_rows * P / dt, "unit": "pairs/s", "cores": threads or os.cpu_count(),
             "kind": job.kind, "seconds_whole_job": dt, "seconds_by_stage": job.seconds,
             "scores_per_sec": cfg["N"] * n_rows / max(job.seconds["score"], 1e-9),
             "sample": "one WHOLE job, not a sample: %s -- calculate_sig_scores on all %d signatures "
                       "(single-threaded Python/NumPy, as the reference is), SignatureScores.ranks, "
                       "sigs_vs_projections over all %d projections; BLAS threads = %s of %d host cores"
                       % ("the unmodified reference (baseline/_ref)" if job.kind == "reference"
                          else "oracle port (row-blocked N x N kernel)", n_rows, P, threads, os.cpu_count() or 0)},
            job)


def cpu_sampled(cfg, wl, budget_s=25.0):
    """For shapes whose whole job takes the host hours (C3): the oracle port on a bounded sample,
    scaled linearly -- scoring on S signatures of the full matrices; rankdata on the same;
    consistency: ONE projection, a block of rows of the N x N kernel against ALL K+R rank columns
    (rows are independent and equal in cost); np.median on 64 columns."""
    import numpy as np
    from oracle import fp_oracle as O, cpu_job as J
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    sample = wl["real"][:12] + wl["rnd"][::max(1, len(wl["rnd"]) // 36)][:36]
    gene_pos = dict((g, i) for i, g in enumerate(wl["genes"]))
    rows = sorted(set(gene_pos[g] for s in sample for g in s.sig_dict if g in gene_pos))
    import torch
    idx = torch.tensor(rows, dtype=torch.long, device=wl["x"].device)
    # only the gene rows the sampled signatures use cross to the host (the per-signature cost of
    # the reference's evaluator does not depend on the rows it does not touch, beyond sig_indices)
    xh, wh = wl["x"].index_select(0, idx).cpu().numpy(), wl["w"].index_select(0, idx).cpu().numpy()
    odata = O.ExpressionMatrix(xh, wh, [wl["genes"][r] for r in rows], wl["cells"])
    to_o = lambda s: O.Signature(s.sig_dict, s.signed, s.source, s.name)
    with J.blas_threads() as threads:
        t0 = time.perf_counter()
        res = O.calculate_sig_scores(odata, [to_o(s) for s in sample], "weighted_avg", None, 5)
        t_score = (time.perf_counter() - t0) / max(1, len(res))
        n_rows = len(wl["real"]) + len(wl["rnd"])
        t0 = time.perf_counter()
        rk = np.array([res[k].ranks for k in res])
        t_rank = (time.perf_counter() - t0) / max(1, len(res))
        reps = (n_rows + rk.shape[0] - 1) // rk.shape[0]
        ranks = np.ascontiguousarray(np.tile(rk, (reps, 1))[:n_rows].T)        # N x (K+R)
        coords = wl["projections"][sorted(wl["projections"].keys())[0]]
        B = 256
        t0 = time.perf_counter()
        rows_done = 0
        while rows_done < N and (rows_done == 0 or (time.perf_counter() - t0) < budget_s * 0.6):
            lo, hi = rows_done, min(N, rows_done + B)
            wt = O.neighbourhood_weights(coords, SIGMA, slice(lo, hi))
            np.abs(ranks[lo:hi] - np.dot(wt, ranks))
            rows_done = hi
        t_cons_rows = (time.perf_counter() - t0) / rows_done
        t0 = time.perf_counter()
        cols = min(n_rows, 16)
        np.median(np.abs(np.random.RandomState(0).normal(size=(N, cols))), axis=0)
        t_med = (time.perf_counter() - t0) / cols
    t_full = n_rows * (t_score + t_rank) + P * (N * t_cons_rows + n_rows * t_med)
    return {"value": n_rows * P / t_full, "unit": "pairs/s", "cores": threads or os.cpu_count(), "kind": "port",
            "scores_per_sec": N / t_score, "extrapolated_s_per_step": t_full,
            "sample": "EXTRAPOLATED (the whole job would take the host %.0f s): oracle port, scoring + rankdata on "
                      "%d signatures; consistency on %d of %d rows of 1 of %d projections against all %d "
                      "signature columns; np.median on %d columns; scaled linearly"
                      % (t_full, len(res), rows_done, N, P, n_rows, cols)}


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores.  Rank 0 only.
    The K timed steps are K consecutive slices of ONE whole job (a step is 1/K of the job); the W
    warm-up steps run slices of a small job of the same kind."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from fastproject_b200 import synth
    from oracle import cpu_job as J
    cfg = WORKLOADS[args.workload]
    G, N, P = cfg["G"], cfg["N"], cfg["P"]
    x, w, z, genes, cells = synth.make_expression(G, N)
    del z
    projections, _ = synth.make_projections(N, P, seed=SEED + 1)
    real, rnd = make_signatures(cfg, genes)
    whole_feasible = args.workload in ("C2", "tiny")
    if not whole_feasible:
        wl = dict(genes=genes, cells=cells, projections=projections, real=real, rnd=rnd)
        emit_reference(args, cfg, real, rnd, None, sampled=(x, w, wl))
        return
    with J.blas_threads() as threads:
        if args.warmup:
            n_small = min(N, 2048)
            small_proj = dict((k, synth.unit_radius(v[:, :n_small].T)) for k, v in list(projections.items())[:2])
            warm = J.CpuJob(np.ascontiguousarray(x[:, :n_small]), np.ascontiguousarray(w[:, :n_small]), genes,
                            cells[:n_small], small_proj, real[:40], rnd[:200], SIGMA, 5, split_projections=True)
            for lo, hi in warm.step_bounds(args.warmup):
                warm.run(lo, hi)
            del warm
        job = J.CpuJob(x, w, genes, cells, projections, real, rnd, SIGMA, 5, split_projections=True)
        t0 = time.perf_counter()
        for lo, hi in job.step_bounds(args.steps):
            job.run(lo, hi)
        dt = time.perf_counter() - t0
    emit_reference(args, cfg, real, rnd, dict(dt=dt, job=job, threads=threads))


def emit_reference(args, cfg, real, rnd, whole, sampled=None):
    P, N = cfg["P"], cfg["N"]
    if whole is not None:
        job, dt = whole["job"], whole["dt"]
        n_rows = len(job.real) + len(job.rnd)
        value = n_rows * P / dt
        base = {"value": value, "unit": "pairs/s", "cores": whole["threads"] or os.cpu_count(), "kind": job.kind,
                "seconds_whole_job": dt, "seconds_by_stage": job.seconds,
                "scores_per_sec": N * n_rows / max(job.seconds["score"], 1e-9),
                "sample": "%d steps = %d consecutive slices of ONE whole job (%d units: calculate_sig_scores on "
                          "chunks of 50 signatures + their ranks, then sigs_vs_projections per projection), %s; "
                          "BLAS threads = %s of %d host cores; scoring is single-threaded Python/NumPy as in the "
                          "reference" % (args.steps, args.steps, len(job.units),
                                         "the unmodified reference (baseline/_ref)" if job.kind == "reference"
                                         else "oracle port (row-blocked N x N kernel)",
                                         whole["threads"], os.cpu_count() or 0)}
        ms_step = 1e3 * dt / args.steps
    else:
        x, w, wl = sampled
        import torch
        wl = dict(wl, x=torch.from_numpy(x), w=torch.from_numpy(w))
        base = cpu_sampled(cfg, wl, budget_s=20.0 * max(1, args.steps) / 3)
        n_rows = len(real) + len(rnd)
        value = base["value"]
        ms_step = 1e3 * n_rows * P / value
    precision = args.precision or "tc"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": describe_config(args.workload, cfg, len(real), len(rnd), precision, args.gpus),
           "cpu_baseline": base,
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
