"""The oracle (oracle/fp_oracle.py) against golden vectors produced by the
live reference (tests/golden/make_golden.py).  CPU only."""
import io
import contextlib
import random

import numpy as np
import pytest

from oracle import fp_oracle as O
from tests import _golden as G

CASES = ["case_synth_odd", "case_synth_even", "case_fixture"]


@pytest.mark.parametrize("case_name", CASES)
@pytest.mark.parametrize("method", G.METHODS)
def test_scores_and_ranks_bit_exact(case_name, method):
    c = G.load(case_name)
    data = G.expression(c, O.ExpressionMatrix)
    prefixes = ["real", "rnd"] + (["rnds"] if "rnds_sig_ptr" in c else [])
    for tag in prefixes:
        sigs = G.unpack_sigs(c, tag + "_", O.Signature)
        res = O.calculate_sig_scores(data, sigs, method, c["Z"], int(c["min_genes"]))
        names = [str(s) for s in c["%s_%s_names" % (tag, method)]]
        assert list(res.keys()) == names
        got = np.array([res[k].scores for k in names]).reshape(len(names), -1)
        assert np.array_equal(got, c["%s_%s_scores" % (tag, method)])
        assert [res[k].numGenes for k in names] == list(c["%s_%s_numgenes" % (tag, method)])
        ranks = np.array([res[k].ranks for k in names]).reshape(len(names), -1)
        assert np.array_equal(ranks, c["%s_%s_ranks" % (tag, method)])


def test_dropped_signatures():
    c = G.load("case_synth_odd")
    names = [str(s) for s in c["real_weighted_avg_names"]]
    assert "TOO_SMALL" not in names and "NO_MATCH" not in names
    assert "EDGE_VALUES" in names
    k = names.index("EDGE_VALUES")
    assert c["real_weighted_avg_numgenes"][k] == 5   # value 0 counts, 0.5 and unknown genes do not


def _svp(c, tag, real_extra=None, **kw):
    data = G.expression(c, O.ExpressionMatrix)
    real = O.calculate_sig_scores(data, G.unpack_sigs(c, "real_", O.Signature), "weighted_avg",
                                  c["Z"], int(c["min_genes"]))
    rnd = O.calculate_sig_scores(data, G.unpack_sigs(c, "rnd_", O.Signature), "weighted_avg",
                                 c["Z"], int(c["min_genes"]))
    if real_extra:
        real.update(real_extra)
    return O.sigs_vs_projections(G.projections(c), real, rnd, **kw)


def _check_svp(out, c, tag, tol=1e-12):
    rows, cols, cons, logq, rcons, rkeys = out
    assert rows == [str(s) for s in c[tag + "_rows"]]
    assert cols == [str(s) for s in c[tag + "_cols"]]
    assert rkeys == [str(s) for s in c[tag + "_rkeys"]]
    np.testing.assert_allclose(cons, c[tag + "_cons"], rtol=tol, atol=0)
    np.testing.assert_allclose(rcons, c[tag + "_rcons"], rtol=tol, atol=0)
    np.testing.assert_allclose(logq, c[tag + "_logq"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("case_name", CASES)
def test_sigs_vs_projections(case_name):
    c = G.load(case_name)
    _check_svp(_svp(c, "svp"), c, "svp")
    if "svp_wide_cons" in c:
        _check_svp(_svp(c, "svp_wide", NEIGHBORHOOD_SIZE=0.8), c, "svp_wide")


@pytest.mark.parametrize("case_name", CASES)
def test_row_blocked_restatement_equals_unblocked(case_name):
    c = G.load(case_name)
    _check_svp(_svp(c, "svp", row_block=17), c, "svp", tol=1e-11)


def test_precomputed_and_factor_branches():
    c = G.load("case_synth_odd")
    cells = [str(x) for x in c["cells"]]
    pre = dict()
    for name in [str(s) for s in c["pre_names"]]:
        v = c["pre_" + name]
        is_factor = v.dtype.kind in "USi"
        vals = v if not is_factor else ([int(x) for x in v] if v.dtype.kind == "i" else [str(x) for x in v])
        pre[name] = O.SignatureScores(vals, name, cells, is_factor, True, 0)
    O._perm_cache = np.zeros((0, 0))
    out = _svp(c, "svpmix", real_extra=pre)
    # factor levels are ints so that list(set(levels)) (Signatures.py:378) has
    # a hash-seed independent order; everything is then reproducible
    _check_svp(out, c, "svpmix")


def test_generate_random_sigs_rng_streams():
    c = G.load("case_randsigs")
    genes = [str(g) for g in c["genes"]]
    random.seed(424242)
    np.random.seed(171717)
    a = O.generate_random_sigs(genes, False, None, 3)
    b = O.generate_random_sigs(genes, True, [7, 13], 5)
    for got, tag in ((a, "a_"), (b, "b_")):
        want = G.unpack_sigs(c, tag, O.Signature)
        assert [s.name for s in got] == [s.name for s in want]
        for s, t in zip(got, want):
            assert list(s.sig_dict.items()) == list(t.sig_dict.items())


def test_p_to_q():
    c = G.load("case_randsigs")
    assert np.array_equal(O.p_to_q(c["p_in"].copy()), c["q_out"])


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_normalization_methods_bit_exact(tag):
    """Oracle restatement of NormalizationMethods.py against outputs of the live reference."""
    c = G.load("case_norm")
    x = c[tag + "_X"]
    assert np.array_equal(O.row_normalization(x.copy()), c[tag + "_rows"])
    assert np.array_equal(O.col_normalization(x.copy()), c[tag + "_cols"])
    assert np.array_equal(O.row_and_col_normalization(x.copy()), c[tag + "_rows_cols"])
    assert np.array_equal(O.col_rank_normalization(x.copy()), c[tag + "_col_rank"])


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_compute_weights_bit_exact(tag):
    """Oracle restatement of Transforms.compute_weights / estimate_non_detection against the
    live reference (never-detected gene -> NaN row, always-detected gene, general L and S)."""
    c = G.load("case_weights")
    x = c[tag + "_X"]
    p_nd = O.estimate_non_detection(x)
    assert np.array_equal(p_nd, c[tag + "_p_nd"])
    w = O.compute_weights(O.logistic_fnr, c[tag + "_params"], x, p_nd)
    assert np.array_equal(w, c[tag + "_weights"], equal_nan=True)
    assert np.isnan(w[2]).all() and (w[4] == 1.0).all()
