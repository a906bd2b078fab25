# -*- coding: utf-8 -*-
"""The step of ``FastProject/Pipelines.py`` right after the hot path: pruning of the
signature x projection matrices to the significant signatures (``Analysis``, :299-349).
"Next" row N4 of SURVEY.md 8f; host code (the matrices are signatures x projections).

Parity: pinned.  In the reference this is an inline block of ``Analysis()``; it runs on this
image's pandas once ``Series[int]`` on a label index means "by position" again (what it meant
when the block was written).  ``tests/config1/run_config1.py golden`` runs the UNMODIFIED
reference pipeline with that shim and records what enters and leaves the block; the goldens
``tests/golden/case_filter_small.npz`` (260 samples, 262 signatures: the -1.3 threshold raised to
the 200th best value) and ``case_filter_large.npz`` (2 050 samples: the 200 best signatures) are
reproduced exactly by the functions below (``tests/test_host_logic.py``).
"""
from __future__ import annotations

import numpy as np

OUTPUT_SIGNATURE_LIMIT = 200
SIGNIFICANCE_THRESHOLD = -1.3          # log10(q) (:321)


def signature_significance(projection_data):
    """Best (smallest) log10 q of every signature over all projections of all filters
    (:304-309).  Returns (keys sorted as the reference's outer-joined frame sorts them, values)."""
    best = dict()
    for proj_data in projection_data:
        p = np.asarray(proj_data["sigProjMatrix_p"], dtype=np.float64)
        for key, row in zip(proj_data["signatureKeys"], p):
            m = np.nan if np.all(np.isnan(row)) else np.nanmin(row)         # DataFrame.min skips NaN
            prev = best.get(key, np.nan)
            best[key] = m if np.isnan(prev) or m < prev else prev            # and so does the outer min
    keys = sorted(best)
    return keys, np.array([best[k] for k in keys], dtype=np.float64)


def significance_threshold(values, n_samples, all_sigs=False):
    """:311-328.  More than 2000 samples: the 200 best signatures; else log10 q <= -1.3, raised
    to the 200th best value when fewer than 200 pass."""
    if all_sigs:
        return 1e99
    limit = min(OUTPUT_SIGNATURE_LIMIT, values.size)
    order = np.argsort(values)                       # NaN last, like Series.argsort's ordering
    nth = values[order[limit - 1]]
    if n_samples > 2000:
        return nth
    return nth if nth > SIGNIFICANCE_THRESHOLD else SIGNIFICANCE_THRESHOLD


def filter_significant_signatures(model, n_samples, all_sigs=False):
    """Prune every ``projData`` of ``model`` (``sigProjMatrix``, ``sigProjMatrix_p``,
    ``signatureKeys``) to the signatures that are precomputed or at least as significant as the
    threshold (:330-349).  Returns the keep map ``name -> bool``.  Like the reference, the
    model's ``signatureScores`` dictionary itself is left as it is (:338-344 rebind a local)."""
    keys, values = signature_significance(model["projectionData"])
    if not keys:
        return dict()
    threshold = significance_threshold(values, n_samples, all_sigs)
    value_of = dict(zip(keys, values))
    keep = dict()
    for name, sig_score in model["signatureScores"].items():
        keep[name] = not (not sig_score.isPrecomputed and value_of[name] > threshold)
    for proj_data in model["projectionData"]:
        mask = np.array([keep[sig] for sig in proj_data["signatureKeys"]], dtype=bool)
        proj_data["sigProjMatrix"] = np.asarray(proj_data["sigProjMatrix"])[mask, :]
        proj_data["sigProjMatrix_p"] = np.asarray(proj_data["sigProjMatrix_p"])[mask, :]
        proj_data["signatureKeys"] = [x for x in proj_data["signatureKeys"] if keep[x]]
    return keep
