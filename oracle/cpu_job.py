# -*- coding: utf-8 -*-
"""One whole job of the hot path on the host CPU, cut into units.  TEST INFRASTRUCTURE ONLY.

Used by ``bench.py`` (``cpu_baseline`` of the GPU arm's line, and ``--impl reference``) and by
the tests.  The job is what the reference pipeline does on this path
(/root/reference/FastProject/Pipelines.py:204, 229, 253): ``calculate_sig_scores`` for the real
and the random signatures, their ranks, ``sigs_vs_projections``.

Two implementations:
  kind "reference"  the UNMODIFIED reference installed in baseline/_ref (oracle/reference.py):
                    its own ``Signatures.calculate_sig_scores`` on chunks of the signature list,
                    ``SignatureScores.ranks``, and ``Signatures.sigs_vs_projections`` -- one call
                    over all projections (exact reference output, used for parity), or one call
                    per projection when the job has to be dealt to many timed steps;
  kind "port"       the restatement in oracle/fp_oracle.py with the N x N kernel evaluated in row
                    blocks (for hosts / shapes where 4 N x N float64 arrays do not fit, and when
                    baseline/_ref is absent).

The job is a list of units executed in order; ``run(lo, hi)`` executes units [lo, hi), so a timed
"step" of the reference arm is 1/steps of ONE whole job and the steps together are the job: no
sample is scaled up.  BLAS threads are set explicitly (torchrun exports OMP_NUM_THREADS=1) and
the number actually in force is reported.
"""
from __future__ import annotations

import contextlib
import io
import os
import time

import numpy as np

from . import fp_oracle as O
from . import reference as R


class _Data(object):
    """Attribute contract of DataTypes.ExpressionData that the evaluators read."""

    def __init__(self, base, weights, row_labels, col_labels):
        self.base = base
        self.weights = weights
        self.row_labels = list(row_labels)
        self.col_labels = list(col_labels)


@contextlib.contextmanager
def blas_threads(n=None):
    """All host cores for BLAS, whatever OMP_NUM_THREADS says; yields the thread count in force."""
    n = n or os.cpu_count() or 1
    try:
        import threadpoolctl
    except Exception:                               # pragma: no cover
        yield None
        return
    with threadpoolctl.threadpool_limits(limits=n):
        counts = [int(p.get("num_threads", 0)) for p in threadpoolctl.threadpool_info()
                  if p.get("user_api") == "blas"]
        yield (max(counts) if counts else None)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class CpuJob(object):
    def __init__(self, x, w, genes, cells, projections, real_sigs, rnd_sigs, sigma=0.33, min_genes=5,
                 kind="auto", split_projections=False, chunk=50, row_block=512):
        ref = R.load() if kind in ("auto", "reference") else None
        if kind == "reference" and ref is None:
            raise RuntimeError("baseline/_ref is not installed")
        self.kind = "reference" if ref is not None else "port"
        self.sigma = sigma
        self.min_genes = min_genes
        self.projections = projections
        self.proj_names = sorted(projections.keys())
        self.n_cells = len(cells)
        self.split_projections = split_projections
        self.row_block = row_block
        if self.kind == "reference":
            self.sig_mod, self.ssm_mod = ref
            mk = self.sig_mod.Signature
        else:
            self.sig_mod = O
            mk = O.Signature
        self.data = _Data(x, w, genes, cells)
        self.real_sigs = [mk(s.sig_dict, s.signed, s.source, s.name) for s in real_sigs]
        self.rnd_sigs = [mk(s.sig_dict, s.signed, s.source, s.name) for s in rnd_sigs]
        self.real, self.rnd = dict(), dict()
        self.cons = self.logq = self.rnd_cons = None
        self.row_names = self.rnd_names = None
        self.seconds = dict(score=0.0, rank=0.0, consistency=0.0)
        self.units = []
        for sigs, dst in ((self.real_sigs, self.real), (self.rnd_sigs, self.rnd)):
            for lo in range(0, len(sigs), chunk):
                self.units.append(("score", self._score_unit, (sigs[lo:lo + chunk], dst)))
        if self.kind == "reference":
            if split_projections:
                for p in self.proj_names:
                    self.units.append(("consistency", self._ref_svp_unit, ([p],)))
            else:
                self.units.append(("consistency", self._ref_svp_unit, (self.proj_names,)))
        else:
            self.units.append(("consistency", self._port_begin, ()))
            for pi in range(len(self.proj_names)):
                for lo in range(0, self.n_cells, row_block):
                    self.units.append(("consistency", self._port_rows, (pi, lo, min(self.n_cells, lo + row_block))))
                self.units.append(("consistency", self._port_finish_projection, (pi,)))
            self.units.append(("consistency", self._port_end, ()))

    # ------------------------------------------------------------------ units
    def _score_unit(self, sigs, dst):
        t0 = time.perf_counter()
        res = _quiet(self.sig_mod.calculate_sig_scores, self.data, sigs, "weighted_avg", None, self.min_genes)
        t1 = time.perf_counter()
        for ss in res.values():
            ss.ranks                                   # scipy rankdata, cached (SigScoreMethods.py:170-178)
        dst.update(res)
        self.seconds["score"] += t1 - t0
        self.seconds["rank"] += time.perf_counter() - t1

    def _ref_svp_unit(self, names):
        t0 = time.perf_counter()
        proj = dict((p, self.projections[p]) for p in names)
        out = _quiet(self.sig_mod.sigs_vs_projections, proj, self.real, self.rnd, self.sigma)
        self.seconds["consistency"] += time.perf_counter() - t0
        rows, cols, cons, logq, rnd_cons, rnd_names = out
        if self.cons is None:
            self.row_names, self.rnd_names = rows, rnd_names
            self.cons = np.zeros((len(rows), len(self.proj_names)))
            self.logq = np.full((len(rows), len(self.proj_names)), np.nan)
            self.rnd_cons = np.zeros((len(rnd_names), len(self.proj_names)))
        idx = [self.proj_names.index(c) for c in cols]
        self.cons[:, idx] = cons
        self.rnd_cons[:, idx] = rnd_cons
        if len(names) == len(self.proj_names):          # Benjamini-Hochberg runs over the whole matrix
            self.logq[:, idx] = logq

    def _port_begin(self):
        t0 = time.perf_counter()
        self.row_names = list(self.real.keys())
        self.rnd_names = list(self.rnd.keys())
        n = self.n_cells
        self._ranks = np.zeros((n, len(self.row_names)))
        for j, k in enumerate(self.row_names):
            self._ranks[:, j] = self.real[k].ranks
        self._rnd_ranks = np.zeros((n, len(self.rnd_names)))
        for j, k in enumerate(self.rnd_names):
            self._rnd_ranks[:, j] = self.rnd[k].ranks
        self._dis = np.empty_like(self._ranks)
        self._rdis = np.empty_like(self._rnd_ranks)
        self.cons = np.zeros((len(self.row_names), len(self.proj_names)))
        self._pval = np.zeros_like(self.cons)
        self.rnd_cons = np.zeros((len(self.rnd_names), len(self.proj_names)))
        self.seconds["consistency"] += time.perf_counter() - t0

    def _port_rows(self, pi, lo, hi):
        t0 = time.perf_counter()
        sl = slice(lo, hi)
        wt = O.neighbourhood_weights(self.projections[self.proj_names[pi]], self.sigma, sl)
        self._dis[sl] = np.abs(self._ranks[sl] - np.dot(wt, self._ranks))
        self._rdis[sl] = np.abs(self._rnd_ranks[sl] - np.dot(wt, self._rnd_ranks))
        self.seconds["consistency"] += time.perf_counter() - t0

    def _port_finish_projection(self, pi):
        t0 = time.perf_counter()
        n = self.n_cells
        med = np.median(self._dis, axis=0)
        rmed = np.median(self._rdis, axis=0)
        table = O.background_table(rmed, [self.rnd[k].numGenes for k in self.rnd_names])
        self._pval[:, pi] = O.background_pvalues(med, [self.real[k].numGenes for k in self.row_names], table)
        self.cons[:, pi] = 1 - med / n
        self.rnd_cons[:, pi] = 1 - rmed / n
        self.seconds["consistency"] += time.perf_counter() - t0

    def _port_end(self):
        q = O.p_to_q(self._pval)
        q[q == 0] = 1e-300
        with np.errstate(invalid="ignore", divide="ignore"):
            self.logq = np.log10(q)
        del self._dis, self._rdis

    # ------------------------------------------------------------------ driving
    def run(self, lo=0, hi=None):
        hi = len(self.units) if hi is None else hi
        for _, fn, args in self.units[lo:hi]:
            fn(*args)

    def step_bounds(self, steps):
        """Unit ranges of ``steps`` consecutive slices that together are the whole job."""
        n = len(self.units)
        return [(s * n // steps, (s + 1) * n // steps) for s in range(steps)]

    def scores_matrix(self, which="real"):
        d = self.real if which == "real" else self.rnd
        return np.array([d[k].scores for k in d]).reshape(len(d), -1)
