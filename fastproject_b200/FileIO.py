# -*- coding: utf-8 -*-
"""B200-side mirror of the writers that put the hot path's results on disk
(``FastProject/FileIO.py``: ``write_matrix`` :154-163, ``write_signature_scores`` :188-201,
the three files of ``write_models`` :272-338 that hold them: ``SignatureScores.txt``,
``ConsistencyMatrix.txt``, ``PMatrix.txt``).  "Next" row N4 of SURVEY.md 8f: formats either side
of the path, host code.

Same file format byte for byte: tab separated, a leading tab before the column labels, every
number as ``"{:.5f}"`` with trailing zeros and a trailing point removed.  The reference formats
number by number in a Python loop (about a second per million values); here a whole row is
formatted by one ``%`` operation and the zeros are stripped by one regular expression, which is
what makes writing 500 signatures x 20 000 cells (10 M numbers) practical.  Scores are read
through ``SignatureScores.scores`` -- device-resident batches are copied to the host once.
"""
from __future__ import annotations

import os
import re

import numpy as np

# "{:.5f}".format(x).rstrip('0').rstrip('.') applied to every field of a formatted row: the zeros
# (and then the point) that end a field, i.e. stand right before a tab or the end of the row
_STRIP = re.compile(r"\.?0+(?=\t|$)")


def _format_row(values):
    """One row of numbers -> tab-joined text, each field as the reference writes it."""
    values = np.asarray(values, dtype=np.float64).ravel()
    if values.size == 0:
        return ""
    text = ("\t".join(["%.5f"] * values.size)) % tuple(values.tolist())
    # "%.5f" always writes the point for finite values and "nan" / "inf" hold no zeros, so a run
    # of zeros right before a tab (or the end) is always the tail of a fraction
    return _STRIP.sub("", text)


def write_matrix(filename, data, row_labels, col_labels):
    """FileIO.py:154-163."""
    data = np.asarray(data)
    with open(filename, "w") as ff:
        ff.write("\t" + "\t".join(col_labels) + "\n")      # extra tab so col labels line up with data
        for i in range(data.shape[0]):
            ff.write(row_labels[i] + "\t")
            ff.write(_format_row(data[i, :]))
            ff.write("\n")


def write_signature_scores(filename, sig_scores_dict, col_labels):
    """FileIO.py:188-201: one row per signature; factor signatures are written as their labels."""
    with open(filename, "w") as ff:
        ff.write("\t" + "\t".join(col_labels) + "\n")
        for name in sig_scores_dict:
            sig_scores = sig_scores_dict[name]
            if sig_scores.isFactor:
                row = "\t".join([name] + list(sig_scores.scores))
            else:
                row = name + "\t" + _format_row(sig_scores.scores)
            ff.write(row + "\n")


def write_path_results(directory, signature_scores, sample_labels, sig_proj_matrix, sig_proj_matrix_p,
                       signature_keys, projection_keys):
    """The three result files of ``write_models`` (:272-338) that come out of the hot path, in
    ``directory``: SignatureScores.txt, ConsistencyMatrix.txt, PMatrix.txt."""
    try:
        os.makedirs(directory)
    except OSError:
        pass
    write_signature_scores(os.path.join(directory, "SignatureScores.txt"), signature_scores, sample_labels)
    write_matrix(os.path.join(directory, "ConsistencyMatrix.txt"), sig_proj_matrix, signature_keys, projection_keys)
    write_matrix(os.path.join(directory, "PMatrix.txt"), sig_proj_matrix_p, signature_keys, projection_keys)
