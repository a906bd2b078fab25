# -*- coding: utf-8 -*-
"""BASELINE config 1 end to end: the reference's own ``fastproject`` command-line pipeline
(``CLI.loadFilesFromDisk -> Pipelines.Analysis -> FileIO.saveResultstoDisk``,
/root/reference/FastProject/CLI.py:223-236) on synthetic files, run twice in separate processes:

  stock        the installed reference (baseline/_ref) as it is;
  accelerated  the same pipeline with the hot-path entry points bound to fastproject_b200
               (the import lines of INTEGRATION.md section 1).

and a comparison of the result files the path feeds (ConsistencyMatrix.txt, PMatrix.txt,
SignatureScores.txt of every filter / projection family).  TEST INFRASTRUCTURE: it imports the
reference and is never imported by the product.

The reference pipeline is bit-rotted on this image's Python / pandas; the
shims below restore the meaning its code had when it was written, without touching its files:
  * ``open(f, 'rU')`` -> ``open(f, 'r')``                       (FileIO.py:78,99; Signatures.py:70,207,566)
  * ``Index & Index`` = set intersection                       (Transforms.py:72)
  * ``Series[int]`` on a label index = by position            (Pipelines.py:318,324)
  * ``DataFrame.iteritems`` = ``items``                        (Interactive.py:286-308)

    python tests/config1/run_config1.py make  DIR [--cells 2000 --genes 5000 --sigs 100]
    python tests/config1/run_config1.py stock DIR
    python tests/config1/run_config1.py accel DIR
    python tests/config1/run_config1.py compare DIR      -> JSON summary on stdout
    python tests/config1/run_config1.py all DIR          (make, both runs as subprocesses, compare)
    python tests/config1/run_config1.py golden DIR OUT.npz   (stock run recording the significance filter)
"""
from __future__ import annotations

import builtins
import contextlib
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- shims
def install_shims():
    import numpy as np
    import pandas as pd
    _open = builtins.open

    def open_compat(file, mode="r", *a, **k):
        if isinstance(mode, str) and "U" in mode:
            mode = mode.replace("U", "") or "r"
        return _open(file, mode, *a, **k)

    builtins.open = open_compat
    io.open = open_compat

    def index_and(self, other):
        return self.intersection(other)

    pd.Index.__and__ = index_and
    if not hasattr(pd.DataFrame, "iteritems"):
        pd.DataFrame.iteritems = pd.DataFrame.items
    if not hasattr(pd.Series, "iteritems"):
        pd.Series.iteritems = pd.Series.items
    _getitem = pd.Series.__getitem__

    def getitem_positional(self, key):
        if isinstance(key, (int, np.integer)) and not isinstance(key, bool) and \
                not pd.api.types.is_integer_dtype(self.index.dtype):
            return self.iloc[int(key)]
        return _getitem(self, key)

    pd.Series.__getitem__ = getitem_positional


# ----------------------------------------------------------------------------- inputs
def make_inputs(work, n_cells=2000, n_genes=5000, n_sigs=100, seed=1335607828):
    """Expression matrix (tab separated, genes x cells, raw counts so that the loader's log
    transform applies) and a signature file in the reference's .txt format."""
    import numpy as np
    os.makedirs(work, exist_ok=True)
    rng = np.random.RandomState(seed)
    # latent 2-D structure + per-gene loadings so that projections and signatures have signal
    n_blobs = 5
    centres = rng.normal(scale=2.0, size=(n_blobs, 2))
    which = rng.randint(0, n_blobs, size=n_cells)
    latent = centres[which] + rng.normal(scale=0.5, size=(n_cells, 2))
    load = rng.normal(scale=0.6, size=(n_genes, 2)) * (rng.uniform(size=(n_genes, 1)) < 0.3)
    base = rng.gamma(2.0, 1.0, size=(n_genes, 1))
    mean = np.exp(0.4 * (load @ latent.T)) * base * 4.0
    counts = rng.poisson(mean).astype(np.float64)
    depth = rng.uniform(0.3, 1.0, size=(1, n_cells))                       # cell quality -> drop-outs
    counts *= (rng.uniform(size=counts.shape) < (0.35 + 0.6 * depth * (mean / (mean + 2.0))))
    genes = ["GENE%04d" % i for i in range(n_genes)]
    cells = ["cell%04d" % i for i in range(n_cells)]
    with open(os.path.join(work, "expression.txt"), "w") as fh:
        fh.write("\t" + "\t".join(cells) + "\n")
        for g, row in zip(genes, counts):
            fh.write(g + "\t" + "\t".join("%g" % v for v in row) + "\n")
    with open(os.path.join(work, "signatures.txt"), "w") as fh:
        for k in range(n_sigs):
            size = int(rng.choice([20, 40, 80, 120]))
            if k < n_sigs // 3:                                             # aligned with the latent axes
                axis = k % 2
                order = np.argsort(load[:, axis])
                pos, neg = order[-size // 2:], order[:size // 2]
                for g in pos:
                    fh.write("SIG_%03d\tplus\t%s\n" % (k, genes[g]))
                for g in neg:
                    fh.write("SIG_%03d\tminus\t%s\n" % (k, genes[g]))
            else:
                for g in rng.choice(n_genes, size=size, replace=False):
                    fh.write("SIG_%03d\t%s\n" % (k, genes[g]))
    hk = rng.choice(n_genes, size=60, replace=False)
    with open(os.path.join(work, "housekeeping.txt"), "w") as fh:
        for g in hk:
            fh.write(genes[g] + "\n")


# ----------------------------------------------------------------------------- runs
def bind_accelerated():
    """route the reference pipeline's hot path through fastproject_b200."""
    import __graft_entry__ as ge
    ge.build()
    import FastProject.Signatures as RS
    import FastProject.SigScoreMethods as RM
    from fastproject_b200 import Signatures as BS, SigScoreMethods as BM
    RS.calculate_sig_scores = BS.calculate_sig_scores
    RS.generate_random_sigs = BS.generate_random_sigs
    RS.sigs_vs_projections = BS.sigs_vs_projections
    for name in ("naive_eval_signature", "weighted_eval_signature", "imputed_eval_signature",
                 "nonzero_eval_signature"):
        setattr(RM, name, getattr(BM, name))


def run_pipeline(work, mode, extra=None, capture=None):
    from oracle import reference
    if reference.load() is None:
        raise SystemExit("baseline/_ref is not installed")
    install_shims()
    if mode == "accel":
        bind_accelerated()
    from FastProject import CLI, Pipelines, FileIO
    out_dir = os.path.join(work, "out_" + mode)
    argv = [os.path.join(work, "expression.txt"), "-s", os.path.join(work, "signatures.txt"),
            "-k", os.path.join(work, "housekeeping.txt"), "-o", out_dir, "--lean"] + list(extra or [])
    old_argv = sys.argv
    sys.argv = ["fastproject"] + argv
    t0 = time.time()
    timing = dict()
    try:
        args = CLI.parseFPArgs()
        loaded = CLI.loadFilesFromDisk(args)
        dir_name = CLI.createOutputDirectories(args)
        import FastProject.Signatures as RS
        for fn in ("calculate_sig_scores", "sigs_vs_projections", "generate_random_sigs"):
            setattr(RS, fn, _timed(getattr(RS, fn), timing, fn))
        calls = []
        if capture:
            inner = RS.sigs_vs_projections

            def recording(*a, **k):
                out = inner(*a, **k)
                calls.append(dict(keys=list(out[0]), cols=list(out[1]), cons=out[2].copy(), logq=out[3].copy()))
                return out
            RS.sigs_vs_projections = recording
        t1 = time.time()
        models, qc_info = Pipelines.Analysis(*loaded, args)
        timing["Analysis"] = time.time() - t1
        if capture:
            import numpy as np
            model = models["Expression"]
            blob = dict(n_samples=np.array(loaded[0].shape[1]), n_calls=np.array(len(calls)),
                        sig_names=np.array(list(model["signatureScores"].keys()), dtype=str),
                        sig_precomputed=np.array([bool(v.isPrecomputed) for v in model["signatureScores"].values()]))
            assert len(calls) == len(model["projectionData"])
            for i, (c, pd_) in enumerate(zip(calls, model["projectionData"])):
                blob["before_keys_%d" % i] = np.array(c["keys"], dtype=str)
                blob["before_cons_%d" % i] = c["cons"]
                blob["before_logq_%d" % i] = c["logq"]
                blob["after_keys_%d" % i] = np.array(pd_["signatureKeys"], dtype=str)
                blob["after_cons_%d" % i] = np.asarray(pd_["sigProjMatrix"])
                blob["after_logq_%d" % i] = np.asarray(pd_["sigProjMatrix_p"])
            np.savez_compressed(capture, **blob)
        FileIO.saveResultstoDisk(models, loaded[1], qc_info, dir_name)
    finally:
        sys.argv = old_argv
    timing["total"] = time.time() - t0
    with open(os.path.join(work, "timing_%s.json" % mode), "w") as fh:
        json.dump(timing, fh)
    return out_dir


def _timed(fn, sink, key):
    def wrapper(*a, **k):
        t0 = time.time()
        try:
            return fn(*a, **k)
        finally:
            sink[key] = sink.get(key, 0.0) + time.time() - t0
    wrapper.__name__ = getattr(fn, "__name__", key)
    return wrapper


# ----------------------------------------------------------------------------- comparison
def _read_matrix(path):
    import numpy as np
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split("\t")[1:]
        rows, vals = [], []
        for line in fh:
            parts = line.rstrip("\n").split("\t")
            rows.append(parts[0])
            vals.append([float(v) if v not in ("", "nan", "NaN") else np.nan for v in parts[1:]])
    return rows, header, np.array(vals, dtype=np.float64).reshape(len(rows), -1)


def compare(work):
    import numpy as np
    a_dir, b_dir = os.path.join(work, "out_stock"), os.path.join(work, "out_accel")
    report = dict(files=[], ok=True)
    for dirpath, _, files in os.walk(a_dir):
        for f in sorted(files):
            if f not in ("ConsistencyMatrix.txt", "PMatrix.txt", "SignatureScores.txt"):
                continue
            rel = os.path.relpath(os.path.join(dirpath, f), a_dir)
            other = os.path.join(b_dir, rel)
            entry = dict(file=rel)
            if not os.path.exists(other):
                entry["missing_in_accelerated"] = True
                report["ok"] = False
                report["files"].append(entry)
                continue
            ra, ca, va = _read_matrix(os.path.join(dirpath, f))
            rb, cb, vb = _read_matrix(other)
            entry["shape"] = list(va.shape)
            entry["labels_identical"] = bool(ra == rb and ca == cb)
            entry["bytes_identical"] = open(os.path.join(dirpath, f), "rb").read() == open(other, "rb").read()
            if va.shape == vb.shape:
                both = np.isfinite(va) & np.isfinite(vb)
                entry["nan_pattern_identical"] = bool(np.array_equal(np.isnan(va), np.isnan(vb)))
                denom = np.maximum(np.abs(va[both]), 1e-300)
                entry["max_rel_err"] = float((np.abs(va[both] - vb[both]) / denom).max()) if both.any() else 0.0
                entry["max_abs_err"] = float(np.abs(va[both] - vb[both]).max()) if both.any() else 0.0
                if f == "PMatrix.txt":
                    entry["ranking_identical"] = bool(np.array_equal(
                        np.argsort(np.nan_to_num(va, nan=9e9), axis=None, kind="stable"),
                        np.argsort(np.nan_to_num(vb, nan=9e9), axis=None, kind="stable")))
                    entry["significance_calls_identical"] = bool(np.array_equal(va < -1.3, vb < -1.3))
                    entry["significant"] = int((va < -1.3).sum())
                tol = 1e-6
                good = entry["labels_identical"] and entry["nan_pattern_identical"] and \
                    (entry["max_rel_err"] <= tol or entry["max_abs_err"] <= 1e-9) and \
                    entry.get("ranking_identical", True) and entry.get("significance_calls_identical", True)
            else:
                good = False
            entry["ok"] = bool(good)
            report["ok"] = report["ok"] and bool(good)
            report["files"].append(entry)
    for mode in ("stock", "accel"):
        p = os.path.join(work, "timing_%s.json" % mode)
        if os.path.exists(p):
            report["timing_" + mode] = json.load(open(p))
    report["n_files"] = len(report["files"])
    report["ok"] = bool(report["ok"] and report["n_files"] > 0)
    return report


def main(argv):
    cmd, work = argv[0], os.path.abspath(argv[1])
    rest = argv[2:]
    if cmd == "make":
        kw = dict()
        for flag, key in (("--cells", "n_cells"), ("--genes", "n_genes"), ("--sigs", "n_sigs")):
            if flag in rest:
                kw[key] = int(rest[rest.index(flag) + 1])
        make_inputs(work, **kw)
    elif cmd in ("stock", "accel"):
        run_pipeline(work, cmd, rest)
    elif cmd == "golden":
        # stock run that also records the significance filter of Analysis() (Pipelines.py:299-349):
        # the sigs_vs_projections outputs going in, the pruned matrices coming out
        run_pipeline(work, "stock", [a for a in rest[1:]], capture=rest[0])
    elif cmd == "compare":
        print(json.dumps(compare(work)))
    elif cmd == "all":
        main(["make", work] + rest)
        for mode in ("stock", "accel"):
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), mode, work],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            with open(os.path.join(work, "log_%s.txt" % mode), "w") as fh:
                fh.write(proc.stdout)
            if proc.returncode != 0:
                sys.stderr.write(proc.stdout[-3000:])
                raise SystemExit("%s run failed" % mode)
        print(json.dumps(compare(work)))
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main(sys.argv[1:])
