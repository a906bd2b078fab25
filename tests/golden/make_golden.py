# -*- coding: utf-8 -*-
"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the authoring container only (needs the read-only reference mount):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the unmodified reference (``/root/reference/FastProject``), runs
its hot-path functions (SigScoreMethods.*_eval_signature, SignatureScores.ranks,
Signatures.calculate_sig_scores / generate_random_sigs / sigs_vs_projections /
p_to_q) on seeded inputs and on the reference's own Tests/TestFiles fixtures,
and stores inputs + outputs as .npz.  The .npz files are committed; this
script is the provenance record.  Nothing in tests/ reads /root/reference.
"""
import io
import os
import random
import sys
import contextlib

sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)

import numpy as np
import scipy

from FastProject import SigScoreMethods as RefSSM   # noqa: E402
from FastProject import Signatures as RefSig        # noqa: E402
from FastProject.DataTypes import ExpressionData    # noqa: E402
from FastProject.Global import RANDOM_SEED          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from fastproject_b200 import synth                  # noqa: E402

METHODS = ["naive", "weighted_avg", "imputed", "only_nonzero"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def pack_sigs(sigs):
    ptr, genes, vals, names, signed = [0], [], [], [], []
    for s in sigs:
        for g, v in s.sig_dict.items():
            genes.append(g)
            vals.append(float(v))
        ptr.append(len(genes))
        names.append(s.name)
        signed.append(bool(s.signed))
    return dict(sig_ptr=np.array(ptr, dtype=np.int64),
                sig_gene=np.array(genes, dtype=str),
                sig_val=np.array(vals, dtype=np.float64),
                sig_name=np.array(names, dtype=str),
                sig_signed=np.array(signed, dtype=bool))


def prefixed(prefix, d):
    return dict((prefix + k, v) for k, v in d.items())


def score_block(edata, sigs, zeros, min_genes, tag):
    out = dict()
    for m in METHODS:
        res = quiet(RefSig.calculate_sig_scores, edata, sigs, m, zeros, min_genes)
        names = list(res.keys())
        out["%s_%s_names" % (tag, m)] = np.array(names, dtype=str)
        out["%s_%s_scores" % (tag, m)] = np.array([res[k].scores for k in names]).reshape(len(names), -1)
        out["%s_%s_numgenes" % (tag, m)] = np.array([res[k].numGenes for k in names], dtype=np.int64)
        out["%s_%s_ranks" % (tag, m)] = np.array([res[k].ranks for k in names]).reshape(len(names), -1)
    return out


def svp_block(projections, real, rnd, tag, nbhd=None):
    if nbhd is None:
        r = quiet(RefSig.sigs_vs_projections, projections, real, rnd)
    else:
        r = quiet(RefSig.sigs_vs_projections, projections, real, rnd, nbhd)
    rows, cols, cons, logq, rcons, rkeys = r
    return {tag + "_rows": np.array(rows, dtype=str),
            tag + "_cols": np.array(cols, dtype=str),
            tag + "_cons": cons, tag + "_logq": logq,
            tag + "_rcons": rcons, tag + "_rkeys": np.array(rkeys, dtype=str)}


def versions():
    return dict(numpy_version=np.array(np.__version__),
                scipy_version=np.array(scipy.__version__),
                reference=np.array("YosefLab/FastProject v1.1.4 @ /root/reference"))


def synth_case(n_genes, n_cells, n_proj, seed, with_precomputed, dup_cells):
    x, w, z, genes, cells = synth.make_expression(n_genes, n_cells, seed=seed)
    if dup_cells:
        # identical cells -> identical scores -> tied ranks
        x[:, 5] = x[:, 3]; w[:, 5] = w[:, 3]; z[:, 5] = z[:, 3]
        x[:, 9] = x[:, 3]; w[:, 9] = w[:, 3]; z[:, 9] = z[:, 3]
    # a cell whose weights vanish on every gene of signature SIG_0000 -> zero normaliser
    sig_defs = synth.make_signature_dicts(genes, 12, seed=seed + 2, n_planted=2)
    for g in sig_defs[0][1]:
        w[genes.index(g), 7] = 0.0
    edata = ExpressionData(x, genes, cells)
    edata.weights = w
    sigs = [RefSig.Signature(d, sg, "synthetic", nm) for nm, d, sg in sig_defs]
    # edge cases of sig_indices: value 0 counts as +1, 0.5 is ignored,
    # genes absent from the matrix are ignored, too-small and empty signatures
    sigs.append(RefSig.Signature({genes[1]: 0, genes[2]: 0.5, genes[3]: -1, genes[4]: 1,
                                  genes[10]: 1, genes[11]: -1, "NOT_A_GENE": 1},
                                 True, "synthetic", "EDGE_VALUES"))
    sigs.append(RefSig.Signature({genes[20]: 1, genes[21]: 1}, False, "synthetic", "TOO_SMALL"))
    sigs.append(RefSig.Signature({"NOPE1": 1, "NOPE2": -1}, True, "synthetic", "NO_MATCH"))

    random.seed(seed)
    np.random.seed(seed)
    rnd_unsigned = RefSig.generate_random_sigs(genes, False, [5, 10, 20], 20)
    rnd_signed = RefSig.generate_random_sigs(genes, True, [5, 12], 8)

    out = dict(X=x, W=w, Z=z, genes=np.array(genes, dtype=str), cells=np.array(cells, dtype=str),
               seed=np.array(seed), min_genes=np.array(5))
    out.update(prefixed("real_", pack_sigs(sigs)))
    out.update(prefixed("rnd_", pack_sigs(rnd_unsigned)))
    out.update(prefixed("rnds_", pack_sigs(rnd_signed)))
    out.update(score_block(edata, sigs, z, 5, "real"))
    out.update(score_block(edata, rnd_unsigned, z, 5, "rnd"))
    out.update(score_block(edata, rnd_signed, z, 5, "rnds"))
    # F-ordered base must give the same answer
    edata_f = ExpressionData(np.asfortranarray(x), genes, cells)
    edata_f.weights = np.asfortranarray(w)
    res_f = quiet(RefSig.calculate_sig_scores, edata_f, sigs, "weighted_avg", z, 5)
    out["real_weighted_avg_scores_forder"] = np.array([res_f[k].scores for k in res_f])

    projections, _ = synth.make_projections(n_cells, n_proj, seed=seed + 1)
    # an isolated cell: every kernel weight of its row underflows to 0
    first = sorted(projections.keys())[0]
    projections[first][:, 4] = [60.0, -45.0]
    out["proj_names"] = np.array(sorted(projections.keys()), dtype=str)
    out["proj_coords"] = np.array([projections[k] for k in sorted(projections.keys())])

    real = quiet(RefSig.calculate_sig_scores, edata, sigs, "weighted_avg", z, 5)
    rnd = quiet(RefSig.calculate_sig_scores, edata, rnd_unsigned, "weighted_avg", z, 5)
    out.update(svp_block(projections, real, rnd, "svp"))
    out.update(svp_block(projections, real, rnd, "svp_wide", 0.8))

    if with_precomputed:
        rng = np.random.RandomState(seed + 9)
        pre = dict()
        time_like = np.repeat([1.0, 2.0, 4.0, 6.0], (n_cells + 3) // 4)[:n_cells]
        pre["TIME_LIKE"] = RefSSM.SignatureScores(time_like, "TIME_LIKE", cells, False, True, 0)
        pre["CONSTANT"] = RefSSM.SignatureScores(np.ones(n_cells) * 3.0, "CONSTANT", cells, False, True, 0)
        pre["SMOOTH"] = RefSSM.SignatureScores(projections[first][0] + 0.1 * rng.normal(size=n_cells),
                                               "SMOOTH", cells, False, True, 0)
        # integer levels: list(set(..)) order is then independent of PYTHONHASHSEED
        levels = [10, 20, 30]
        fac = [levels[(i * 7 // 3) % 3] for i in range(n_cells)]
        pre["STIM"] = RefSSM.SignatureScores(fac, "STIM", cells, True, True, 0)
        pre["ONE_LEVEL"] = RefSSM.SignatureScores(["A"] * n_cells, "ONE_LEVEL", cells, True, True, 0)
        out["pre_names"] = np.array(list(pre.keys()), dtype=str)
        out["pre_TIME_LIKE"] = time_like
        out["pre_CONSTANT"] = pre["CONSTANT"].scores
        out["pre_SMOOTH"] = pre["SMOOTH"].scores
        out["pre_STIM"] = np.array(fac, dtype=np.int64)
        out["pre_ONE_LEVEL"] = np.array(pre["ONE_LEVEL"].scores, dtype=str)
        mixed = dict(real)
        mixed.update(pre)
        RefSig._bg_dist = np.zeros((0, 0))
        out.update(svp_block(projections, mixed, rnd, "svpmix"))
    out.update(versions())
    return out


def fixture_case():
    """The reference's own Tests/TestFiles inputs (smallData 1038 x 65,
    weights.txt, sigsSmall.gmt, input_projection_good.txt)."""
    tf = os.path.join(REF, "FastProject", "Tests", "TestFiles")

    def read_matrix(path):
        with open(path) as fh:
            cols = fh.readline().rstrip("\n").split("\t")[1:]
            rows, vals = [], []
            for line in fh:
                parts = line.rstrip("\n").split("\t")
                if len(parts) < 2:
                    continue
                rows.append(parts[0].upper())
                vals.append([float(v) for v in parts[1:]])
        return np.array(vals), rows, cols

    x0, genes, cells = read_matrix(os.path.join(tf, "smallData.txt"))
    w, wgenes, _ = read_matrix(os.path.join(tf, "weights.txt"))
    assert wgenes == genes
    z = (x0 == 0).astype(np.float64)
    x = synth.zscore_rows(x0)
    edata = ExpressionData(x, genes, cells)
    edata.weights = w
    sigs = []
    with open(os.path.join(tf, "sigsSmall.gmt")) as fh:
        for line in fh:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 3:
                continue
            name = parts[0]
            d = dict()
            signed = False
            for tok in parts[2:]:
                if tok == "":
                    continue
                d[tok.upper()] = 1
            sigs.append(RefSig.Signature(d, signed, "sigsSmall.gmt", name))
    # signed variants: second half of each gene list negative (GMT _plus/_minus pairs
    # are merged by the reference parser; here we only need Signature objects)
    for s in list(sigs[:4]):
        d = dict()
        for i, g in enumerate(s.sig_dict):
            d[g] = 1 if i % 2 == 0 else -1
        sigs.append(RefSig.Signature(d, True, "sigsSmall.gmt", s.name + "_SIGNED"))
    random.seed(RANDOM_SEED)
    np.random.seed(RANDOM_SEED)
    rnd = RefSig.generate_random_sigs(genes, False, [5, 10, 20, 50, 100], 12)

    out = dict(X=x, W=w, Z=z, genes=np.array(genes, dtype=str), cells=np.array(cells, dtype=str),
               min_genes=np.array(5))
    out.update(prefixed("real_", pack_sigs(sigs)))
    out.update(prefixed("rnd_", pack_sigs(rnd)))
    out.update(score_block(edata, sigs, z, 5, "real"))
    out.update(score_block(edata, rnd, z, 5, "rnd"))

    xy = []
    with open(os.path.join(tf, "input_projection_good.txt")) as fh:
        for line in fh:
            p = line.split()
            if len(p) == 3:
                xy.append([float(p[1]), float(p[2])])
    xy = np.array(xy)
    assert xy.shape[0] == len(cells)
    projections = {"Input": synth.unit_radius(xy)}
    more, _ = synth.make_projections(len(cells), 2, seed=77)
    projections.update(more)
    out["proj_names"] = np.array(sorted(projections.keys()), dtype=str)
    out["proj_coords"] = np.array([projections[k] for k in sorted(projections.keys())])
    real = quiet(RefSig.calculate_sig_scores, edata, sigs, "weighted_avg", z, 5)
    rsc = quiet(RefSig.calculate_sig_scores, edata, rnd, "weighted_avg", z, 5)
    out.update(svp_block(projections, real, rsc, "svp"))
    out.update(versions())
    return out


def randsig_case():
    """RNG-stream parity of generate_random_sigs (two global streams)."""
    genes = ["F%04d" % i for i in range(500)]
    out = dict(genes=np.array(genes, dtype=str))
    random.seed(424242)
    np.random.seed(171717)
    a = RefSig.generate_random_sigs(genes, False, None, 3)
    b = RefSig.generate_random_sigs(genes, True, [7, 13], 5)
    out.update(prefixed("a_", pack_sigs(a)))
    out.update(prefixed("b_", pack_sigs(b)))
    # p_to_q on a matrix with ties and exact zeros
    rng = np.random.RandomState(5)
    p = rng.uniform(size=(9, 4))
    p[2, 1] = 0.0
    p[3, 3] = p[0, 0]
    p[4, 2] = 1.0
    out["p_in"] = p
    out["q_out"] = RefSig.p_to_q(p.copy())
    out.update(versions())
    return out


def norm_case():
    """NormalizationMethods.py on a small matrix with a constant gene, a constant cell, ties,
    and a long row (pairwise summation splits above 128 elements)."""
    from FastProject import NormalizationMethods as RefNorm
    rng = np.random.RandomState(RANDOM_SEED % (2 ** 32))
    out = dict(versions())
    for tag, shape in (("small", (23, 37)), ("wide", (9, 1237))):
        x = np.log1p(rng.gamma(0.5, 4.0, size=shape)) * (rng.uniform(size=shape) < 0.6)
        x[3, :] = 2.5                      # constant gene: sigma 0 -> 1
        if tag == "small":
            x[:, 5] = 0.0                  # constant cell
        out[tag + "_X"] = x
        out[tag + "_rows"] = RefNorm.row_normalization(x.copy())
        out[tag + "_cols"] = RefNorm.col_normalization(x.copy())
        out[tag + "_rows_cols"] = RefNorm.row_and_col_normalization(x.copy())
        out[tag + "_col_rank"] = RefNorm.col_rank_normalization(x.copy())
    return out


def weights_case():
    """Transforms.compute_weights (:266-316) of the live reference with the logistic curve of
    create_false_neg_map (:87-88; a closure there, restated here) and seeded per-cell
    parameters; p_nd from Transforms.estimate_non_detection (:341-381)."""
    import pandas as pd
    from FastProject import Transforms as RefT

    def func(xvals, x0, a, L=0, S=1):
        return L + S / (1 + np.exp((xvals - x0) * a))

    rng = np.random.RandomState((RANDOM_SEED + 7) % (2 ** 32))
    out = dict(versions())
    for tag, shape in (("small", (37, 53)), ("wide", (11, 1237))):
        g, n = shape
        x = np.log1p(rng.gamma(0.6, 5.0, size=shape)) * (rng.uniform(size=shape) < 0.55)
        x[2, :] = 0.0                                   # never detected: mu_h = 0/0
        x[4, :] = 1.0 + rng.uniform(size=n)             # always detected: pnd == 0 -> 1/N
        x[6, :] = 0.0
        x[6, 3] = 2.0                                   # detected once
        genes = ["G%03d" % i for i in range(g)]
        cells = ["c%04d" % i for i in range(n)]
        params = np.zeros((4, n))
        params[0] = rng.uniform(0.5, 4.0, size=n)
        params[1] = rng.uniform(0.0, 2.0, size=n)
        params[1, :3] = [0.0, 2.0, 1e-3]
        params[2] = 0.0
        params[3] = 1.0
        if tag == "small":                              # general L, S as the signature allows
            params[2, 10:20] = 0.05
            params[3, 10:20] = 0.9
        data = ExpressionData(x.copy(), genes, cells)
        p_nd = RefT.estimate_non_detection(pd.DataFrame(x, index=genes, columns=cells))
        with np.errstate(all="ignore"):
            w = RefT.compute_weights(func, params, data, p_nd)
        out[tag + "_X"] = x
        out[tag + "_params"] = params
        out[tag + "_p_nd"] = p_nd.values
        out[tag + "_weights"] = w
    return out


def writers_case():
    """FileIO.write_matrix (:154-163) and write_signature_scores (:188-201) of the live reference
    on awkward numbers (trailing zeros, negative zero, tiny, huge, nan, inf, half-way cases)."""
    import tempfile
    from FastProject import FileIO as RefIO
    rng = np.random.RandomState((RANDOM_SEED + 11) % (2 ** 32))
    special = np.array([0.0, -0.0, 1.0, -1.0, 10.0, 100.0, 1000000.0, 0.5, 0.00001, 0.000004, 0.000005,
                        -0.000004, np.nan, np.inf, -np.inf, 1e20, 123456.789, 99.999996, 0.1 + 0.2, 1e-300,
                        5e-6, 1.5e-5, 100.000001, 10.1, 1010.0, 0.100001, 2.5e-6, 0.000015])
    vals = np.concatenate([rng.normal(size=272) * 10.0 ** rng.randint(-7, 8, size=272), special])
    data = vals.reshape(12, 25)
    rows = ["sig %d" % i for i in range(12)]
    cols = ["proj%d" % i for i in range(25)]
    tmp = tempfile.mkdtemp()
    RefIO.write_matrix(os.path.join(tmp, "m.txt"), data, rows, cols)
    scores = dict()
    order = []
    for i in range(5):
        name = "S%d" % i
        order.append(name)
        scores[name] = RefSSM.SignatureScores(list(data[i]), name, cols, isFactor=False, isPrecomputed=False,
                                              numGenes=3)
    scores["F"] = RefSSM.SignatureScores(["a", "b", "c", "a", "b"] * 5, "F", cols, isFactor=True,
                                         isPrecomputed=True, numGenes=0)
    order.append("F")
    RefIO.write_signature_scores(os.path.join(tmp, "s.txt"), scores, cols)
    out = dict(versions())
    out["data"] = data
    out["rows"] = np.array(rows, dtype=str)
    out["cols"] = np.array(cols, dtype=str)
    out["score_order"] = np.array(order, dtype=str)
    out["matrix_text"] = np.array(open(os.path.join(tmp, "m.txt")).read())
    out["scores_text"] = np.array(open(os.path.join(tmp, "s.txt")).read())
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "writers":
        np.savez_compressed(os.path.join(HERE, "case_writers.npz"), **writers_case())
        return
    np.savez_compressed(os.path.join(HERE, "case_writers.npz"), **writers_case())
    if len(sys.argv) > 1 and sys.argv[1] == "weights":
        np.savez_compressed(os.path.join(HERE, "case_weights.npz"), **weights_case())
        return
    np.savez_compressed(os.path.join(HERE, "case_weights.npz"), **weights_case())
    np.savez_compressed(os.path.join(HERE, "case_norm.npz"), **norm_case())
    np.savez_compressed(os.path.join(HERE, "case_synth_odd.npz"),
                        **synth_case(240, 65, 3, 1001, True, False))
    np.savez_compressed(os.path.join(HERE, "case_synth_even.npz"),
                        **synth_case(200, 64, 2, 2002, False, True))
    np.savez_compressed(os.path.join(HERE, "case_fixture.npz"), **fixture_case())
    np.savez_compressed(os.path.join(HERE, "case_randsigs.npz"), **randsig_case())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
