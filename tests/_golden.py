"""Helpers to load tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
METHODS = ["naive", "weighted_avg", "imputed", "only_nonzero"]


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def _value(v):
    # dict values were stored as float; ints compare equal either way but keep
    # the type a parser would have produced
    return int(v) if float(v).is_integer() else float(v)


def unpack_sigs(case, prefix, signature_cls):
    ptr = case[prefix + "sig_ptr"]
    genes = case[prefix + "sig_gene"]
    vals = case[prefix + "sig_val"]
    names = case[prefix + "sig_name"]
    signed = case[prefix + "sig_signed"]
    out = []
    for k in range(len(names)):
        d = dict()
        for j in range(ptr[k], ptr[k + 1]):
            d[str(genes[j])] = _value(vals[j])
        out.append(signature_cls(d, bool(signed[k]), "golden", str(names[k])))
    return out


def expression(case, matrix_cls):
    return matrix_cls(case["X"], case["W"], [str(g) for g in case["genes"]],
                      [str(c) for c in case["cells"]])


def projections(case):
    return dict((str(n), case["proj_coords"][i]) for i, n in enumerate(case["proj_names"]))
