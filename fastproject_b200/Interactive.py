# -*- coding: utf-8 -*-
"""B200 mirror of the scoring wrappers of ``FastProject/Interactive.py`` -- the pandas
facing caller of the hot path.

Same names, arguments and return values as the reference
(/root/reference/FastProject/Interactive.py:143-209 ``calculate_sig_scores``,
:212-326 ``sigs_vs_projection``, :19-21 re-export of ``generate_random_sigs``), written
against current pandas (the reference's ``Index & Index`` and ``DataFrame.iteritems`` no
longer exist).  All arithmetic is delegated to ``fastproject_b200.Signatures``.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import synth
from .SigScoreMethods import SignatureScores
from .Signatures import calculate_sig_scores as _calculate_sig_scores
from .Signatures import generate_random_sigs  # noqa: F401  (re-exported like the reference)
from .Signatures import sigs_vs_projections as _sigs_vs_projections

__all__ = ["generate_random_sigs", "calculate_sig_scores", "sigs_vs_projection"]


def calculate_sig_scores(data, signatures, weights=None, method='weighted_avg',
                         zero_locations=None, min_signature_genes=0):
    """Interactive version of Signatures.calculate_sig_scores that uses DataFrames
    (Interactive.py:143-209).

    :param data: pandas.DataFrame genes x samples
    :param signatures: list of Signature
    :param weights: pandas.DataFrame like ``data`` (all ones if None)
    :return: (scores DataFrame NUM_SAMPLES x NUM_SIGNATURES, num_genes Series)
    """
    w = weights.values if weights is not None else np.ones_like(data.values)
    expression = synth.ExpressionMatrix(data.values, w, list(data.index), list(data.columns))
    result_dict = _calculate_sig_scores(expression, signatures, method=method,
                                        zero_locations=zero_locations,
                                        min_signature_genes=min_signature_genes)
    names = list(result_dict.keys())
    if names:
        first = result_dict[names[0]]
        batch = getattr(first, "_batch", None)
        rows = [result_dict[n]._row for n in names]
        block = batch.host_scores()[rows].T if batch is not None else \
            np.column_stack([result_dict[n].scores for n in names])
        all_scores = pd.DataFrame(block, index=first.sample_labels, columns=names)
    else:
        all_scores = pd.DataFrame(index=list(data.columns))
    num_genes = pd.Series([result_dict[n].numGenes for n in names], index=names)
    return all_scores, num_genes


def _aligned(frame, index, what):
    assert len(index.intersection(frame.index)) == len(index), what + " does not cover the projection's samples"
    return frame.loc[index]


def sigs_vs_projection(projection, sig_scores, num_genes,
                       random_sig_scores, random_num_genes,
                       numerical_sig_scores=None, factor_sig_scores=None,
                       NEIGHBORHOOD_SIZE=0.33):
    """Interactive wrapper for Signatures.sigs_vs_projections (Interactive.py:212-326).

    :param projection: DataFrame NUM_SAMPLES x 2
    :param sig_scores: DataFrame NUM_SAMPLES x NUM_SIGNATURES, ``num_genes`` Series of matches
    :param random_sig_scores / random_num_genes: the same for the background signatures
    :param numerical_sig_scores / factor_sig_scores: optional precomputed signatures
    :return: (consistencies, pvals (log10 q), null_consistencies) DataFrames
    """
    sig_scores = _aligned(sig_scores, projection.index, "sig_scores")
    random_sig_scores = _aligned(random_sig_scores, projection.index, "random_sig_scores")
    if numerical_sig_scores is not None:
        numerical_sig_scores = _aligned(numerical_sig_scores, projection.index, "numerical_sig_scores")
    if factor_sig_scores is not None:
        factor_sig_scores = _aligned(factor_sig_scores, projection.index, "factor_sig_scores")

    projections = {'proj_name': np.ascontiguousarray(projection.values.T, dtype=np.float64)}
    sig_scores_dict = {}
    for name, column in sig_scores.items():
        sig_scores_dict[name] = SignatureScores(column.values, name, column.index.tolist(),
                                                isFactor=False, isPrecomputed=False,
                                                numGenes=num_genes[name])
    if numerical_sig_scores is not None:
        for name, column in numerical_sig_scores.items():
            sig_scores_dict[name] = SignatureScores(column.values, name, column.index.tolist(),
                                                    isFactor=False, isPrecomputed=True, numGenes=-1)
    if factor_sig_scores is not None:
        for name, column in factor_sig_scores.items():
            sig_scores_dict[name] = SignatureScores(column.values.tolist(), name, column.index.tolist(),
                                                    isFactor=True, isPrecomputed=True, numGenes=-1)
    random_sig_scores_dict = {}
    for name, column in random_sig_scores.items():
        random_sig_scores_dict[name] = SignatureScores(column.values, name, column.index.tolist(),
                                                       isFactor=False, isPrecomputed=False,
                                                       numGenes=random_num_genes[name])

    (row_labels, col_labels, sig_proj_matrix, sig_proj_matrix_p, random_sig_proj_matrix,
     random_sig_score_keys) = _sigs_vs_projections(projections, sig_scores_dict, random_sig_scores_dict,
                                                   NEIGHBORHOOD_SIZE=NEIGHBORHOOD_SIZE)
    consistencies = pd.DataFrame(sig_proj_matrix, index=row_labels, columns=col_labels)
    null_consistencies = pd.DataFrame(random_sig_proj_matrix, index=random_sig_score_keys, columns=col_labels)
    pvals = pd.DataFrame(sig_proj_matrix_p, index=row_labels, columns=col_labels)
    return consistencies, pvals, null_consistencies
