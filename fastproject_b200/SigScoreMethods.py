# -*- coding: utf-8 -*-
"""B200 mirror of ``FastProject/SigScoreMethods.py``.

Same names, argument meaning and error behaviour as the reference module
(/root/reference/FastProject/SigScoreMethods.py): the four evaluators are
module-level functions ``f(self, signature, zeros, min_signature_genes)`` whose
``self`` is an ExpressionData-like object (only ``.base``, ``.weights``,
``.row_labels``, ``.col_labels`` are read), returning ``SignatureScores``.
The arithmetic runs in ``fp_score_signatures`` / ``fp_rank_columns``
(include/fastproject_b200.h); scores are bit-identical to the reference.

Each evaluator call scores ONE signature; the batched entry point that keeps
the expression matrix resident and scores thousands of signatures per launch
is ``Signatures.calculate_sig_scores``.
"""
from __future__ import annotations

import numpy as np

from . import _scoring


class SignatureScores(object):
    """Represents a Signature evaluated on a set of samples
    (reference: SigScoreMethods.py:165-206).

    ``scores`` and ``ranks`` may be backed by a device-resident batch
    (``_scoring.ScoreBatch``); the host arrays are then materialised on first
    access.  ``sigs_vs_projections`` consumes the device copy directly.
    """

    def __init__(self, scores, name, sample_labels, isFactor, isPrecomputed, numGenes):
        self._scores = scores
        self.name = name
        self.sample_labels = sample_labels
        self.isFactor = isFactor
        self.isPrecomputed = isPrecomputed
        self._ranks = None
        self.numGenes = numGenes
        self._batch = None
        self._row = -1

    @classmethod
    def _from_batch(cls, batch, row, name, sample_labels, numGenes):
        obj = cls(None, name, sample_labels, False, False, numGenes)
        obj._batch = batch
        obj._row = row
        return obj

    @property
    def scores(self):
        if self._scores is None and self._batch is not None:
            self._scores = self._batch.host_scores()[self._row]
        return self._scores

    @scores.setter
    def scores(self, value):
        self._scores = value
        self._ranks = None
        self._batch = None

    @property
    def ranks(self):
        """scipy.stats.rankdata(scores, method="average") -- SigScoreMethods.py:170-178."""
        if self.isFactor:
            raise Exception("Factor signature scores have no rank")
        if self._ranks is None:
            if self._batch is not None:
                self._ranks = self._batch.host_ranks()[self._row]
            else:
                self._ranks = _scoring.rank_vector(self._scores)
        return self._ranks

    def to_JSON(self):
        """Dictionary of the same fields the reference serialises (:189-206),
        as a JSON string."""
        import json
        out = {"name": self.name, "scores": np.asarray(self.scores).tolist(),
               "isFactor": self.isFactor, "isPrecomputed": self.isPrecomputed}
        if not self.isFactor:
            out["ranks"] = np.asarray(self.ranks).tolist()
        return json.dumps(out)


def _evaluate_one(method, data, signature, zeros, min_signature_genes):
    batch = _scoring.score_batch(data, [signature], method, zeros, min_signature_genes,
                                 raise_on_reject=True)
    return SignatureScores._from_batch(batch, 0, signature.name, data.col_labels,
                                       int(batch.num_genes[0]))


def naive_eval_signature(self, signature, zeros, min_signature_genes):
    """Naive eval signature just sums the columns * sign (all weights = 1).
    Reference: SigScoreMethods.py:18-48."""
    return _evaluate_one("naive", self, signature, zeros, min_signature_genes)


def weighted_eval_signature(self, signature, zeros, min_signature_genes):
    """Weighted average using weights in self.weights.
    Reference: SigScoreMethods.py:51-82."""
    return _evaluate_one("weighted_avg", self, signature, zeros, min_signature_genes)


def imputed_eval_signature(self, signature, zeros, min_signature_genes):
    """Imputes likely values for zeros using the weight.
    Reference: SigScoreMethods.py:85-128."""
    return _evaluate_one("imputed", self, signature, zeros, min_signature_genes)


def nonzero_eval_signature(self, signature, zeros, min_signature_genes):
    """Scores using only nonzero entries.  Reference: SigScoreMethods.py:131-162."""
    return _evaluate_one("only_nonzero", self, signature, zeros, min_signature_genes)
