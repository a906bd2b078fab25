# -*- coding: utf-8 -*-
"""B200 mirror of ``FastProject/NormalizationMethods.py`` -- the step immediately
before the scoring hot path (``data.get_normalized_copy(norm_method)``,
/root/reference/FastProject/Pipelines.py:202).

Same function names and meaning as the reference (NormalizationMethods.py:17-61).
Each function accepts a NumPy array (returns a NumPy array, like the reference) or a
CUDA tensor (returns a CUDA tensor, which ``calculate_sig_scores`` takes as ``data.base``
without another upload).  The arithmetic runs in ``fp_normalize_rows`` /
``fp_normalize_cols`` / ``fp_rank_columns_ex`` and is bit-identical to NumPy / SciPy.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _engine, _native


def _to_dev(data):
    if isinstance(data, torch.Tensor) and data.is_cuda:
        return _engine.row_major(data), True
    return _engine.to_device(np.asarray(data, dtype=np.float64)), False


def _back(dev, was_tensor):
    return dev if was_tensor else _engine.to_host(dev)


def no_normalization(data):
    """Does nothing.  Returns original data (NormalizationMethods.py:17-21)."""
    return data


def col_normalization(data):
    """Perform z-normalization on all columns (NormalizationMethods.py:24-31)."""
    dev, was = _to_dev(data)
    return _back(_engine.normalize_cols(dev), was)


def row_normalization(data):
    """Perform z-normalization on all rows (NormalizationMethods.py:34-41)."""
    dev, was = _to_dev(data)
    return _back(_engine.normalize_rows(dev), was)


def row_and_col_normalization(data):
    """Normalize rows, then columns (NormalizationMethods.py:44-50)."""
    dev, was = _to_dev(data)
    return _back(_engine.normalize_cols(_engine.normalize_rows(dev)), was)


def col_rank_normalization(data):
    """Column-wise ranks instead of values, ties get the smallest rank
    (``rankdata(..., method="min")``, NormalizationMethods.py:53-61)."""
    dev, was = _to_dev(data)
    g, n = dev.shape
    ld = _engine.padded(g)
    cols = torch.zeros((n, ld), dtype=torch.float64, device=dev.device)      # one cell per row
    cols[:, :g] = dev.t()
    ranks = _engine.rank_rows(cols, g, method=_native.RANK_MIN)
    return _back(ranks[:, :g].t().contiguous(), was)
