# -*- coding: utf-8 -*-
"""B200 mirror of the weight-producing part of ``FastProject/Transforms.py`` -- the step that
makes the W matrix the scoring hot path reads (``Pipelines.py:120`` ->
``Transforms.compute_weights``, /root/reference/FastProject/Transforms.py:266-316, with
``estimate_non_detection`` :341-381 as its input).

Same names and argument meaning as the reference.  ``compute_weights`` runs in
``fp_compute_weights`` (one CUDA kernel: detected-mean per gene, logistic false-negative rate
per gene x cell, weights); NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out (which
``calculate_sig_scores`` takes as ``data.weights`` without another upload).

The per-cell curve fit itself (``create_false_neg_map``: scipy L-BFGS-B on 30 quantile points
per cell) stays with the reference; its outputs ``(fit_func, params)`` are what
``compute_weights`` takes.  The kernel evaluates the logistic family
``L + S / (1 + exp((mu - x0) * a))`` that the reference fits: a ``fit_func`` that is not this
family is refused (there is no CPU fallback).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _engine


def logistic_fnr(xvals, x0, a, L=0, S=1):
    """The false-negative-rate curve of ``create_false_neg_map`` (Transforms.py:87-88)."""
    return L + S / (1 + np.exp((xvals - x0) * a))


def estimate_non_detection(data):
    """Probability of non-detection of every measurement: the zero mask
    (Transforms.py:341-381 with ``DO_GMM = False``).  DataFrame in -> DataFrame out, like the
    reference; arrays and CUDA tensors are accepted too.  ``compute_weights(..., p_nd=None)``
    forms the same mask on the device without materialising it."""
    if isinstance(data, torch.Tensor):
        return (data == 0).to(torch.float64)
    if hasattr(data, "values") and hasattr(data, "index"):
        import pandas as pd
        return pd.DataFrame((data.values == 0).astype("float"), index=data.index, columns=data.columns)
    return (np.asarray(data) == 0).astype("float")


_PROBE_X = np.array([-3.0, -0.5, 0.0, 0.7, 2.0, 5.5, 11.0])
_PROBE_PARAMS = [(1.3, 0.8, 0.0, 1.0), (4.0, 2.0, 0.05, 0.9), (0.2, 0.0, 0.0, 1.0)]


def _check_fit_func(fit_func):
    """Accept the reference's logistic closure (or anything numerically equal to it)."""
    if fit_func is None or fit_func is logistic_fnr:
        return
    try:
        for prm in _PROBE_PARAMS:
            got = np.asarray(fit_func(_PROBE_X, *prm), dtype=np.float64).ravel()
            if got.shape != _PROBE_X.shape or not np.allclose(got, logistic_fnr(_PROBE_X, *prm), rtol=1e-12, atol=0):
                raise ValueError
    except Exception:
        raise NotImplementedError("fastproject_b200.Transforms.compute_weights evaluates the logistic "
                                  "false-negative curve L + S/(1 + exp((x - x0)*a)) of "
                                  "Transforms.create_false_neg_map; the given fit_func is a different function")


def compute_weights(fit_func, params, data, p_nd=None):
    """Weights of the data from the false-negative curves (Transforms.py:266-316):
    p(not expressed | not detected) for the non-detected values, 1.0 for detected ones.

    :param fit_func: the function returned by ``create_false_neg_map`` (or ``logistic_fnr`` / None)
    :param params: (4 x Num_Samples) parameters x0, a, L, S of every cell's curve
    :param data: ExpressionData-like (``.base``; ``.row_labels`` / ``.col_labels`` select from a
        DataFrame ``p_nd`` as the reference does), a NumPy array or a CUDA tensor
    :param p_nd: probability of non-detection, same shape (DataFrame, array, tensor), or None
        for the zero mask of ``estimate_non_detection``
    :return: (Num_Genes x Num_Samples) weights; NumPy unless ``data`` was a CUDA tensor
    """
    _check_fit_func(fit_func)
    base = data.base if hasattr(data, "base") and not isinstance(data, (np.ndarray, torch.Tensor)) else data
    as_tensor = isinstance(base, torch.Tensor) and base.is_cuda
    x = _engine.row_major(base) if as_tensor else _engine.expression_cache.get(np.asarray(base))
    prm = params if isinstance(params, torch.Tensor) else np.asarray(params, dtype=np.float64)
    if tuple(prm.shape) != (4, x.shape[1]):
        raise ValueError("params must be 4 x Num_Samples (got %s for %d samples)" % (tuple(prm.shape), x.shape[1]))
    prm = prm.to(device=x.device, dtype=torch.float64) if isinstance(prm, torch.Tensor) else _engine.to_device(prm)
    pnd_dev = None
    if p_nd is not None:
        if hasattr(p_nd, "loc") and hasattr(data, "row_labels"):
            p_nd = p_nd.loc[data.row_labels]
            p_nd = p_nd.loc[:, data.col_labels].values
        elif hasattr(p_nd, "values") and not isinstance(p_nd, torch.Tensor):
            p_nd = p_nd.values
        if isinstance(p_nd, torch.Tensor) and p_nd.is_cuda:
            pnd_dev = _engine.row_major(p_nd)
        else:
            pnd_dev = _engine.to_device(np.asarray(p_nd, dtype=np.float64))
        if tuple(pnd_dev.shape) != tuple(x.shape):
            raise ValueError("p_nd must have the shape of the data")
    w = _engine.compute_weights(x, prm, pnd_dev)
    return w if as_tensor else _engine.to_host(w)
