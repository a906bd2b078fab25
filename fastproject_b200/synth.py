# -*- coding: utf-8 -*-
"""Deterministic synthetic inputs for the scoring core.

Host-side NumPy only; used by ``bench.py`` and the tests to build inputs of
the shapes BASELINE.json names.  Nothing here is on the hot path.

Recipe (seed = Global.RANDOM_SEED unless given):
  X0 = log1p(Gamma(0.5, 4)) * Bernoulli(0.6)      (about 40 % exact zeros)
  zeros = (X0 == 0)
  X  = per-gene z-score of X0        (NormalizationMethods.py:34-41)
  W  = where(X0 > 0, 1, U(0,1))      (shape of Transforms.compute_weights)
  real signatures: size in {20,50,100,150}, half signed, a few planted on a
  2-D gradient; projections: Gaussian-blob mixtures pushed through the
  reference's projection normalisation (Projections.py:100-113).
"""
from __future__ import annotations

import numpy as np

RANDOM_SEED = 1335607828  # Global.py:71


class ExpressionMatrix(object):
    """Plain carrier with the attribute contract of ``ExpressionData``
    (DataTypes.py:16): ``.base`` G x N float64, ``.weights`` G x N,
    ``.row_labels``, ``.col_labels``."""

    def __init__(self, base, weights, row_labels, col_labels):
        self.base = base
        self.weights = weights
        self.row_labels = list(row_labels)
        self.col_labels = list(col_labels)


def zscore_rows(x):
    mu = x.mean(axis=1, keepdims=True)
    sd = x.std(axis=1, keepdims=True)
    sd[sd == 0] = 1.0
    return (x - mu) / sd


def unit_radius(xy):
    """(N,2) -> (2,N): centred, 90th-percentile radius scaled to one."""
    xy = np.array(xy, dtype=np.float64)
    xy -= xy.mean(axis=0, keepdims=True)
    r90 = np.percentile(np.sqrt((xy ** 2).sum(axis=1)), 90)
    if r90 > 0:
        xy /= r90
    return np.ascontiguousarray(xy.T)


def make_expression(n_genes, n_cells, seed=RANDOM_SEED, planted=None,
                    chunk=2048):
    """Returns (X, W, zeros, gene_names, cell_names).  ``planted`` is an
    optional (n_cells,) gradient added to the first 150 genes so that some
    signatures are genuinely consistent with a projection."""
    rng = np.random.RandomState(seed)
    x = np.empty((n_genes, n_cells), dtype=np.float64)
    w = np.empty((n_genes, n_cells), dtype=np.float64)
    z = np.empty((n_genes, n_cells), dtype=np.float64)
    for lo in range(0, n_genes, chunk):
        hi = min(n_genes, lo + chunk)
        raw = np.log1p(rng.gamma(0.5, 4.0, size=(hi - lo, n_cells)))
        if planted is not None and lo < 150:
            k = min(hi, 150) - lo
            raw[:k] += np.maximum(planted, 0)[None, :] * rng.uniform(0.5, 1.5, size=(k, 1))
        raw *= (rng.uniform(size=raw.shape) < 0.6)
        z[lo:hi] = (raw == 0)
        w[lo:hi] = np.where(raw > 0, 1.0, rng.uniform(size=raw.shape))
        x[lo:hi] = zscore_rows(raw)
    genes = ["G%05d" % i for i in range(n_genes)]
    cells = ["C%06d" % i for i in range(n_cells)]
    return x, w, z, genes, cells


def make_projections(n_cells, n_proj, seed=RANDOM_SEED + 1):
    """dict name -> (2,N); also returns a (N,) gradient usable as ``planted``."""
    rng = np.random.RandomState(seed)
    out = dict()
    gradient = None
    for p in range(n_proj):
        n_blobs = rng.randint(3, 9)
        centres = rng.normal(scale=2.0, size=(n_blobs, 2))
        which = rng.randint(0, n_blobs, size=n_cells)
        xy = centres[which] + rng.normal(scale=0.6, size=(n_cells, 2))
        noise = rng.uniform(size=n_cells) < 0.05
        xy[noise] = rng.uniform(-4, 4, size=(int(noise.sum()), 2))
        coords = unit_radius(xy)
        out["Proj%02d" % p] = coords
        if gradient is None:
            gradient = coords[0] + 0.5 * coords[1]
    return out, gradient


def make_signature_dicts(genes, n_sigs, seed=RANDOM_SEED + 2, n_planted=5):
    """list of (name, dict gene->sign, signed) for the 'real' signatures."""
    rng = np.random.RandomState(seed)
    out = []
    n_genes = len(genes)
    for k in range(n_sigs):
        size = int(rng.choice([20, 50, 100, 150]))
        size = min(size, n_genes)
        signed = bool(k % 2)
        if k < n_planted:
            idx = rng.choice(min(150, n_genes), size=min(size, min(150, n_genes)), replace=False)
            signed = False
        else:
            idx = rng.choice(n_genes, size=size, replace=False)
        sg = rng.choice([-1, 1], size=idx.size) if signed else np.ones(idx.size, dtype=int)
        out.append(("SIG_%04d" % k, dict((genes[i], int(s)) for i, s in zip(idx, sg)), signed))
    return out
