"""Parity of the CUDA path (through the C ABI and the Python mirror of the
reference API) with the oracle and with golden vectors produced by the live
reference.  Bit-exact for scores and ranks; 1e-9 relative for consistency
values in FP64 mode (BLAS summation order is not reproducible), 1e-6 relative
is the north-star tolerance."""
import io
import contextlib
import random

import numpy as np
import pytest

from oracle import fp_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu

CASES = ["case_synth_odd", "case_synth_even", "case_fixture"]
CONS_RTOL = 1e-9       # FP64 mode, measured << this; north-star tolerance is 1e-6
TC_CONS_RTOL = 2e-8    # tensor-core mode (exact digit-sliced int8 contraction, 32-bit fixed-point weights)
LOGQ_RTOL = 1e-6
PRECISIONS = ["fp64", "tc"]


def _rtol(precision):
    return CONS_RTOL if precision == "fp64" else TC_CONS_RTOL


def _logq_atol(precision):
    # |d log10 q| = |dq/q| / ln 10: 1e-6 absolute is 2.3e-6 relative on q
    return 1e-9 if precision == "fp64" else 1e-6


@pytest.fixture(scope="module")
def fp():
    import __graft_entry__ as ge
    ge.build()
    from fastproject_b200 import Signatures, SigScoreMethods, synth, _engine
    import types
    return types.SimpleNamespace(S=Signatures, M=SigScoreMethods, synth=synth, E=_engine)


def _data(fp, c):
    return fp.synth.ExpressionMatrix(c["X"], c["W"], [str(g) for g in c["genes"]],
                                     [str(x) for x in c["cells"]])


# ------------------------------------------------------------------ scoring
@pytest.mark.parametrize("case_name", CASES)
@pytest.mark.parametrize("method", G.METHODS)
def test_scores_and_ranks_bit_exact_vs_reference_golden(fp, case_name, method):
    c = G.load(case_name)
    data = _data(fp, c)
    for tag in ["real", "rnd"] + (["rnds"] if "rnds_sig_ptr" in c else []):
        sigs = G.unpack_sigs(c, tag + "_", fp.S.Signature)
        res = fp.S.calculate_sig_scores(data, sigs, method, c["Z"], int(c["min_genes"]))
        names = [str(s) for s in c["%s_%s_names" % (tag, method)]]
        assert list(res.keys()) == names
        got = np.array([res[k].scores for k in names]).reshape(len(names), -1)
        assert np.array_equal(got, c["%s_%s_scores" % (tag, method)]), (tag, method)
        assert [res[k].numGenes for k in names] == list(c["%s_%s_numgenes" % (tag, method)])
        ranks = np.array([res[k].ranks for k in names]).reshape(len(names), -1)
        assert np.array_equal(ranks, c["%s_%s_ranks" % (tag, method)])
        first = res[names[0]]
        assert first.isFactor is False and first.isPrecomputed is False
        assert first.sample_labels == data.col_labels


def test_forder_and_bool_zero_mask_inputs(fp):
    c = G.load("case_synth_odd")
    data = fp.synth.ExpressionMatrix(np.asfortranarray(c["X"]), np.asfortranarray(c["W"]),
                                     [str(g) for g in c["genes"]], [str(x) for x in c["cells"]])
    sigs = G.unpack_sigs(c, "real_", fp.S.Signature)
    res = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", c["Z"], 5)
    got = np.array([res[k].scores for k in res])
    assert np.array_equal(got, c["real_weighted_avg_scores_forder"])
    res = fp.S.calculate_sig_scores(data, sigs, "only_nonzero", c["Z"].astype(bool), 5)
    got = np.array([res[k].scores for k in res])
    assert np.array_equal(got, c["real_only_nonzero_scores"])


def test_single_signature_evaluators_and_errors(fp):
    c = G.load("case_synth_even")
    data = _data(fp, c)
    sigs = G.unpack_sigs(c, "real_", fp.S.Signature)
    fns = {"naive": fp.M.naive_eval_signature, "weighted_avg": fp.M.weighted_eval_signature,
           "imputed": fp.M.imputed_eval_signature, "only_nonzero": fp.M.nonzero_eval_signature}
    for m, fn in fns.items():
        names = [str(s) for s in c["real_%s_names" % m]]
        ss = fn(data, sigs[0], c["Z"], 5)
        k = names.index(sigs[0].name)
        assert np.array_equal(ss.scores, c["real_%s_scores" % m][k])
        assert ss.numGenes == c["real_%s_numgenes" % m][k] and ss.name == sigs[0].name
        small = [s for s in sigs if s.name == "TOO_SMALL"][0]
        with pytest.raises(ValueError, match="Too few genes match signature"):
            fn(data, small, c["Z"], 5)
        nomatch = [s for s in sigs if s.name == "NO_MATCH"][0]
        with pytest.raises(ValueError, match="No genes match signature"):
            fn(data, nomatch, c["Z"], 0)
    # function default min_signature_genes=1e99 discards everything (Signatures.py:631)
    assert fp.S.calculate_sig_scores(data, sigs) == {}
    with pytest.raises(UnboundLocalError):
        fp.S.calculate_sig_scores(data, sigs, "bogus", c["Z"], 0)


def test_scores_bit_exact_vs_oracle_midsize(fp):
    x, w, z, genes, cells = fp.synth.make_expression(1500, 2100, seed=31)
    w[:, 17] = 0.0                                            # zero normaliser for every signature
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    random.seed(5); np.random.seed(5)
    sigs = fp.S.generate_random_sigs(genes, True, [5, 10, 50, 200], 40)
    odata = O.ExpressionMatrix(x, w, genes, cells)
    osigs = [O.Signature(s.sig_dict, s.signed, s.source, s.name) for s in sigs]
    for m in G.METHODS:
        got = fp.S.calculate_sig_scores(data, sigs, m, z, 5)
        want = O.calculate_sig_scores(odata, osigs, m, z, 5)
        assert list(got.keys()) == list(want.keys())
        a = np.array([got[k].scores for k in want])
        b = np.array([want[k].scores for k in want])
        assert np.array_equal(a, b), m
        assert np.array_equal(np.array([got[k].ranks for k in want]),
                              np.array([want[k].ranks for k in want])), m


@pytest.mark.parametrize("method", G.METHODS)
@pytest.mark.parametrize("g,n,k", [(300, 16, 9), (700, 333, 70), (3100, 1000, 600), (1030, 4100, 1100)])
def test_staged_scoring_kernel_equals_gather_kernel(fp, method, g, n, k, monkeypatch):
    """The TMA-staged scoring kernel (gene rows through shared memory, one pass per group of 512
    signatures) and the gather kernel do the same arithmetic in the same order: BIT-IDENTICAL
    scores, every method, odd cell counts (padded row pitch), partial gene blocks, partial
    signature groups, signed and unsigned signatures of every size; and both equal the oracle."""
    x, w, z, genes, cells = fp.synth.make_expression(g, n, seed=g + n)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    rng = np.random.RandomState(k)
    sigs = []
    for i in range(k):
        size = int(rng.choice([1, 2, 5, 17, 64, 200, min(g, 400)]))
        idx = rng.choice(g, size=min(size, g), replace=False)
        sg = rng.choice([-1, 1], size=idx.size) if i % 2 else np.ones(idx.size, dtype=int)
        sigs.append(fp.S.Signature(dict((genes[j], int(v)) for j, v in zip(idx, sg)), bool(i % 2), "t", "S%d" % i))
    out = dict()
    for kernel in ("gather", "staged"):
        monkeypatch.setenv("FASTPROJECT_B200_SCORE", kernel)
        res = fp.S.calculate_sig_scores(data, sigs, method, z, 1)
        assert len(res) == k
        out[kernel] = np.array([res[s.name].scores for s in sigs])
    assert np.array_equal(out["gather"], out["staged"])
    odata = O.ExpressionMatrix(x, w, genes, cells)
    some = sigs[::max(1, k // 25)]
    want = O.calculate_sig_scores(odata, [O.Signature(s.sig_dict, s.signed, "t", s.name) for s in some], method, z, 1)
    for row, s in zip(out["staged"][::max(1, k // 25)], some):
        assert np.array_equal(row, want[s.name].scores), s.name


def test_rank_ties_constant_negative_zero_and_nan(fp):
    from scipy.stats import rankdata
    rng = np.random.RandomState(0)
    rows = [np.round(rng.normal(size=1001), 1), np.ones(1001), rng.normal(size=1001),
            np.r_[0.0, -0.0, 0.0, rng.normal(size=998)], rng.randint(0, 5, size=1001).astype(float)]
    for v in rows:
        assert np.array_equal(fp.M.SignatureScores(v, "v", [], False, True, 0).ranks,
                              rankdata(v, method="average"))
    v = rng.normal(size=77); v[5] = np.nan
    assert np.isnan(fp.M.SignatureScores(v, "v", [], False, True, 0).ranks).all()
    with pytest.raises(Exception, match="Factor signature scores have no rank"):
        fp.M.SignatureScores(["a"], "f", [], True, True, 0).ranks


@pytest.mark.parametrize("path", ["fused", "sort"])
def test_rank_short_keys_hard_cases(fp, path, monkeypatch):
    """The rank path bins the cells by a short monotone key (rows up to 48 000 cells: a counting
    sort inside one CTA, "fused"; longer rows, or FASTPROJECT_B200_RANK=sort: a segmented radix
    sort) and orders the cells sharing a key bin by their doubles: exercise every bin regime
    against scipy (bit-exact), both tie rules, both paths."""
    import torch
    if path == "sort":
        monkeypatch.setenv("FASTPROJECT_B200_RANK", "sort")
    from scipy.stats import rankdata
    from fastproject_b200 import _engine as E, _native
    rng = np.random.RandomState(3)
    n = 5000
    rows = [
        1.0 + np.arange(n) * 1e-13,                                    # distinct doubles, one outlier below:
        rng.permutation(1.0 + np.arange(n) * 1e-13),                   # ... everything in a few key bins
        np.r_[rng.permutation(np.repeat(rng.normal(size=40), 100)), rng.normal(size=n - 4000)],  # long tie runs
        np.r_[np.inf, -np.inf, rng.normal(size=n - 2)],                # infinities -> bit-pattern keys
        np.r_[np.inf, np.inf, -np.inf, np.zeros(n - 3)],
        rng.normal(size=n) * 1e-320,                                   # denormals
        np.r_[1e308, -1e308, rng.normal(size=n - 2)],                  # span overflows
        np.where(rng.rand(n) < 0.6, 0.0, rng.gamma(1.0, size=n)),      # expression-like: a zero bin + a tail
        np.zeros(n),
        rng.randint(0, 2, size=n).astype(float),
        np.r_[np.nan, rng.normal(size=n - 1)],
    ]
    rows[0][0] = -1e6
    rows[1][7] = 1e9
    x = np.ascontiguousarray(np.array(rows))
    dev = torch.as_tensor(x).cuda()
    for method, name in ((_native.RANK_AVERAGE, "average"), (_native.RANK_MIN, "min")):
        got = E.rank_rows(dev, n, method).cpu().numpy()
        for i, v in enumerate(rows):
            if np.isnan(v).any():
                assert np.isnan(got[i]).all()
            else:
                assert np.array_equal(got[i], rankdata(v, method=name)), (i, name)
    # padded leading dimension and many rows
    k, ld = 300, 1056
    y = rng.normal(size=(k, ld)); y[::7, :1001] = np.round(y[::7, :1001], 1)
    got = E.rank_rows(torch.as_tensor(y).cuda(), 1001).cpu().numpy()
    for i in range(k):
        assert np.array_equal(got[i, :1001], rankdata(y[i, :1001]))
    # the largest rows of the fused path, one beyond it, and tiny ones
    for m in (1, 2, 31, 24500, 24501, 47999, 48000, 48001):
        z = rng.normal(size=(3, m)); z[1] = np.round(z[1], 2); z[2, : m // 2] = 0.0
        zd = torch.zeros((3, E.padded(m)), dtype=torch.float64, device="cuda")
        zd[:, :m] = torch.as_tensor(z).cuda()
        got = E.rank_rows(zd, m).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i, :m], rankdata(z[i])), (m, i)


@pytest.mark.parametrize("path", ["fused", "sort"])
def test_rank_large_mixed_bins_are_not_quadratic(fp, path, monkeypatch):
    """Rows whose cells almost all share one short-key bin although their doubles differ (an
    outlier squeezes the scale; small counts next to the zeros of an expression column): the bin
    is re-quantised on its own range, so such a row costs milliseconds, not seconds -- and the
    ranks stay bit-identical to scipy."""
    import time
    import torch
    from scipy.stats import rankdata
    from fastproject_b200 import _engine as E, _native
    if path == "sort":
        monkeypatch.setenv("FASTPROJECT_B200_RANK", "sort")
    rng = np.random.RandomState(5)
    n = 40000
    rows = [
        np.r_[1e12, rng.normal(size=n - 1)],                                     # one outlier, 39 999 distinct values in bin 0
        np.r_[1e12, np.where(rng.rand(n - 1) < 0.6, 0.0, rng.gamma(1.0, size=n - 1))],   # zeros + a tail, squeezed
        np.r_[1e15, rng.randint(0, 50, size=n - 1).astype(float)],               # small counts: 50 tie runs in one bin
        np.r_[np.inf, -np.inf, rng.normal(size=n - 2) * 1e-300],                 # bit-pattern keys at both levels
    ]
    x = np.ascontiguousarray(np.array(rows))
    dev = torch.zeros((len(rows), E.padded(n)), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.as_tensor(x).cuda()
    for method, name in ((_native.RANK_AVERAGE, "average"), (_native.RANK_MIN, "min")):
        E.rank_rows(dev, n, method)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = E.rank_rows(dev, n, method)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        got = got.cpu().numpy()
        for i, v in enumerate(rows):
            assert np.array_equal(got[i, :n], rankdata(v, method=name)), (i, name)
        assert dt < 0.05, "ranking %d squeezed rows took %.3f s" % (len(rows), dt)


def test_small_uploads_bypass_copy_engine_bit_exact(fp):
    """Per-call tables go host -> device through fp_pull_host (a kernel reading pinned memory):
    every dtype / odd size must arrive byte for byte, also while a large copy is in flight."""
    import torch
    from fastproject_b200 import _engine as E
    rng = np.random.RandomState(11)
    big_host = torch.empty((64 << 20,), dtype=torch.uint8, pin_memory=True)
    big_dev = torch.empty_like(big_host, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        big_dev.copy_(big_host, non_blocking=True)                  # keeps the copy engine busy
    cases = [rng.normal(size=n) for n in (0, 1, 2, 3, 17, 4099, 200001)]
    cases += [rng.randint(-100, 100, size=n).astype(np.int8) for n in (1, 15, 16, 17, 33333)]
    cases += [rng.randint(0, 1 << 30, size=(7, 913)).astype(np.int32), rng.normal(size=(3, 5, 7)),
              np.asfortranarray(rng.normal(size=(50, 60))), rng.uniform(size=1000) < 0.5]
    for a in cases:
        want = np.asarray(a, dtype=np.float64) if a.dtype == np.bool_ else a
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int8): torch.int8,
               np.dtype(np.int32): torch.int32}[want.dtype]
        got = E.to_device(a, tdt).cpu().numpy()
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want)
    # staging buffers are recycled only after the kernel that reads them has finished
    outs = [(v, E.to_device(v)) for v in (rng.normal(size=50000) for _ in range(40))]
    for v, d in outs:
        assert np.array_equal(d.cpu().numpy(), v)
    torch.cuda.synchronize()


# ------------------------------------------------------------------ kernels through the engine
@pytest.mark.parametrize("n", [1, 2, 7, 64, 65, 1000, 4097])
def test_row_medians_exact(fp, n):
    import torch
    rng = np.random.RandomState(n)
    rows = np.vstack([np.abs(rng.normal(size=n)), rng.normal(size=n) * 1e3,
                      rng.randint(0, 3, size=n).astype(float), np.zeros(n),
                      np.r_[rng.uniform(size=n - 1), np.inf][:n] if n > 1 else np.ones(1)])
    ld = fp.E.padded(n)
    dev = torch.zeros((rows.shape[0], ld), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(rows).cuda()
    got = fp.E.row_medians(dev, n).cpu().numpy()
    assert np.array_equal(got, np.median(rows, axis=1))
    if n > 2:
        rows[1, n // 2] = np.nan
        dev[:, :n] = torch.from_numpy(rows).cuda()
        got = fp.E.row_medians(dev, n).cpu().numpy()
        assert np.isnan(got[1]) and np.array_equal(got[[0, 2, 3]], np.median(rows[[0, 2, 3]], axis=1))


def test_background_pvalues_vs_oracle(fp):
    import torch
    rng = np.random.RandomState(9)
    rnd_ng = [int(v) for v in np.repeat([5, 10, 20, 50, 100, 200], 500)]
    rng.shuffle(rnd_ng)
    rmed = rng.normal(loc=500, scale=4, size=len(rnd_ng))
    real_ng = [int(v) for v in rng.randint(1, 300, size=64)]
    med = rng.normal(loc=470, scale=30, size=64)
    table = O.background_table(rmed, rnd_ng)
    want = O.background_pvalues(med, real_ng, table)
    order, ptr, grp = fp.S._background_groups(rnd_ng, real_ng)
    t = lambda a, dt: torch.from_numpy(np.asarray(a)).to("cuda", dt)
    p, mu, sd = fp.E.background_pvalues(t(med, torch.float64), t(grp, torch.int32), t(rmed, torch.float64),
                                        t(order, torch.int32), t(ptr, torch.int32))
    assert np.array_equal(mu.cpu().numpy(), table[:, 1])      # NumPy pairwise order reproduced
    assert np.array_equal(sd.cpu().numpy(), table[:, 2])
    np.testing.assert_allclose(p.cpu().numpy(), want, rtol=1e-12, atol=0)


def test_background_stats_tree_shapes(fp):
    """Group sizes on every side of the pairwise tree's thresholds (short loop below 8, one leaf up
    to 128, splits rounded to multiples of 8): mean and std bit-identical to NumPy."""
    import torch
    rng = np.random.RandomState(10)
    sizes = [1, 2, 7, 8, 9, 127, 128, 129, 255, 1000, 4099, 20011]
    rnd_ng = [int(v) for v in np.repeat(np.arange(1, len(sizes) + 1), sizes)]
    rng.shuffle(rnd_ng)
    rmed = rng.normal(loc=500, scale=4, size=len(rnd_ng))
    real_ng = [int(v) for v in rng.randint(1, len(sizes) + 1, size=40)]
    med = rng.normal(loc=497, scale=5, size=40)
    table = O.background_table(rmed, rnd_ng)
    want = O.background_pvalues(med, real_ng, table)
    order, ptr, grp = fp.S._background_groups(rnd_ng, real_ng)
    t = lambda a, dt: torch.from_numpy(np.asarray(a)).to("cuda", dt)
    p, mu, sd = fp.E.background_pvalues(t(med, torch.float64), t(grp, torch.int32), t(rmed, torch.float64),
                                        t(order, torch.int32), t(ptr, torch.int32))
    assert np.array_equal(mu.cpu().numpy(), table[:, 1])
    assert np.array_equal(sd.cpu().numpy(), table[:, 2])
    np.testing.assert_allclose(p.cpu().numpy(), want, rtol=1e-12, atol=0, equal_nan=True)


# ------------------------------------------------------------------ consistency
def _svp(fp, c, extra=None, **kw):
    data = _data(fp, c)
    real = fp.S.calculate_sig_scores(data, G.unpack_sigs(c, "real_", fp.S.Signature), "weighted_avg",
                                     c["Z"], int(c["min_genes"]))
    rnd = fp.S.calculate_sig_scores(data, G.unpack_sigs(c, "rnd_", fp.S.Signature), "weighted_avg",
                                    c["Z"], int(c["min_genes"]))
    if extra:
        real.update(extra)
    return fp.S.sigs_vs_projections(G.projections(c), real, rnd, **kw)


def _check_svp(out, c, tag, rtol=CONS_RTOL):
    logq_atol = 1e-9 if rtol == CONS_RTOL else 1e-6
    rows, cols, cons, logq, rcons, rkeys = out
    assert rows == [str(s) for s in c[tag + "_rows"]]
    assert cols == [str(s) for s in c[tag + "_cols"]]
    assert rkeys == [str(s) for s in c[tag + "_rkeys"]]
    np.testing.assert_allclose(cons, c[tag + "_cons"], rtol=rtol, atol=0)
    np.testing.assert_allclose(rcons, c[tag + "_rcons"], rtol=rtol, atol=0)
    np.testing.assert_allclose(logq, c[tag + "_logq"], rtol=LOGQ_RTOL, atol=logq_atol)
    # identical ranking of the signature x projection pairs and significance calls
    assert np.array_equal(np.argsort(logq, axis=None, kind="stable"),
                          np.argsort(c[tag + "_logq"], axis=None, kind="stable"))
    assert np.array_equal(logq < -1.3, c[tag + "_logq"] < -1.3)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("case_name", CASES)
def test_sigs_vs_projections_vs_reference_golden(fp, case_name, precision):
    c = G.load(case_name)
    _check_svp(_svp(fp, c, precision=precision), c, "svp", rtol=_rtol(precision))
    if "svp_wide_cons" in c:
        _check_svp(_svp(fp, c, NEIGHBORHOOD_SIZE=0.8, precision=precision), c, "svp_wide",
                   rtol=_rtol(precision))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_precomputed_numeric_and_factor_branches(fp, precision):
    c = G.load("case_synth_odd")
    cells = [str(x) for x in c["cells"]]
    pre = dict()
    for name in [str(s) for s in c["pre_names"]]:
        v = c["pre_" + name]
        is_factor = v.dtype.kind in "USi"
        vals = v if not is_factor else ([int(x) for x in v] if v.dtype.kind == "i" else [str(x) for x in v])
        pre[name] = fp.M.SignatureScores(vals, name, cells, is_factor, True, 0)
    fp.S._bg_dist = np.zeros((0, 0))
    _check_svp(_svp(fp, c, extra=pre, precision=precision), c, "svpmix", rtol=_rtol(precision))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_per_cell_sigma_extension_reduces_to_scalar(fp, precision):
    c = G.load("case_synth_even")
    n = c["X"].shape[1]
    out = _svp(fp, c, sigma_per_cell=np.full(n, 0.33), precision=precision)
    _check_svp(out, c, "svp", rtol=_rtol(precision))


def _tc_case(fp, n, k, seed, outliers=0, dup=0, per_cell=False, binary=False, kind=0):
    """Same inputs through the FP64 verifier kernel and the tensor-core kernel."""
    import torch
    from scipy.stats import rankdata
    rng = np.random.RandomState(seed)
    proj, _ = fp.synth.make_projections(n, 1, seed=seed)
    coords = proj["Proj00"].copy()
    for o in range(outliers):
        coords[:, o] = [6.0 + 3 * o, -5.0 - o]
    for d in range(dup):
        coords[:, n - 1 - d] = coords[:, d + 5]
    if binary:
        vals = (rng.uniform(size=(k, n)) < 0.3).astype(np.float64)
    else:
        vals = np.array([rankdata(rng.normal(size=n) + (coords[0] if r % 3 == 0 else 0)) for r in range(k)])
        vals[k - 1] = rankdata(np.round(rng.normal(size=n), 1))      # many ties -> half-integer ranks
    ld = fp.E.padded(n)
    dev = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(vals).cuda()
    c = torch.from_numpy(coords).cuda()
    sig = 0.33 ** 2
    if per_cell:
        sig = torch.from_numpy((0.2 + 0.3 * rng.uniform(size=n)) ** 2).cuda()
    a = fp.E.consistency(c, sig, dev, n, precision="fp64", output_kind=kind)[:, :n].cpu().numpy()
    b = fp.E.consistency(c, sig, dev, n, precision="tc", output_kind=kind)[:, :n].cpu().numpy()
    return a, b


@pytest.mark.parametrize("n,k,kw", [
    (1, 2, {}), (2, 3, {}), (127, 1, {}), (129, 127, {}), (300, 128, dict(dup=3)),
    (1000, 5, dict(outliers=2)), (4097, 260, dict(outliers=2, dup=3)), (2000, 40, dict(per_cell=True)),
    (33000, 3, dict(outliers=1)),      # three value digits, three j chunks
])
def test_tensor_core_kernel_matches_fp64_kernel(fp, n, k, kw):
    a, b = _tc_case(fp, n, k, 31 + n, **kw)
    assert not np.isnan(b).any()
    # absolute error of the prediction, in rank units, relative to the rank range
    assert np.abs(a - b).max() <= 1e-7 * max(n, 16)
    np.testing.assert_allclose(np.median(b, axis=1), np.median(a, axis=1), rtol=0, atol=1e-8 * max(n, 16))


@pytest.mark.parametrize("n,k,budget", [(1000, 260, 1 << 20), (4097, 130, 5 << 20), (33000, 3, 40 << 20)])
def test_tensor_core_weight_slabs_do_not_change_the_result(fp, n, k, budget, monkeypatch):
    """The weight digit planes are generated slab by slab (cells i) when N x N x 4 bytes exceeds
    the budget (100 000+ cells).  Shrinking the budget runs the many-slab path at test sizes:
    the output must be BIT-IDENTICAL to the one-slab run (same arithmetic, different staging)."""
    a, one = _tc_case(fp, n, k, 77 + n, outliers=1)
    monkeypatch.setenv("FASTPROJECT_B200_SLAB_BYTES", str(budget))
    _, many = _tc_case(fp, n, k, 77 + n, outliers=1)
    assert np.array_equal(one, many)
    assert np.abs(a - one).max() <= 1e-7 * max(n, 16)


@pytest.mark.parametrize("layout", ["blobs", "outliers", "duplicates", "line", "one_point_cloud"])
def test_nearest_neighbour_grid_equals_brute_force(fp, layout, monkeypatch):
    """The row scale of the tensor-core path is the nearest-neighbour distance of every cell.
    From 4 096 cells on it comes from a uniform-grid search; it must be the brute-force kernel's
    minimum exactly, so the whole output is BIT-IDENTICAL between the two."""
    import torch
    from scipy.stats import rankdata
    rng = np.random.RandomState(21)
    n, k = 6000, 5
    if layout == "blobs":
        xy = np.concatenate([rng.normal(c, s, size=(n // 4, 2)) for c, s in ((-1, .02), (0, .3), (.4, .05), (2, .2))])
    elif layout == "outliers":
        xy = rng.normal(0, 0.3, size=(n, 2)); xy[0] = [40.0, -35.0]; xy[1] = [-90.0, 2.0]
    elif layout == "duplicates":
        xy = rng.normal(0, 0.3, size=(n, 2)); xy[100:700] = xy[5]; xy[700:900] = xy[6]
    elif layout == "line":
        xy = np.stack([np.linspace(-1, 1, n), np.zeros(n)], axis=1)
    else:
        xy = np.tile([[0.25, -0.5]], (n, 1)); xy[:50] += rng.normal(0, 1e-3, size=(50, 2))
    coords = torch.from_numpy(np.ascontiguousarray(xy.T)).cuda()
    ld = fp.E.padded(n)
    vals = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    vals[:, :n] = torch.from_numpy(np.array([rankdata(rng.normal(size=n)) for _ in range(k)])).cuda()
    grid = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="tc")[:, :n].cpu().numpy()
    monkeypatch.setenv("FASTPROJECT_B200_NN", "brute")
    brute = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="tc")[:, :n].cpu().numpy()
    assert np.array_equal(grid, brute, equal_nan=True)
    monkeypatch.delenv("FASTPROJECT_B200_NN")
    ref = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="fp64")[:, :n].cpu().numpy()
    assert np.abs(ref - grid).max() <= 1e-7 * n


@pytest.mark.parametrize("n,k,kind", [(6000, 300, 0), (20000, 40, 0), (33000, 5, 0), (6000, 20, 1)])
def test_z_order_and_plane_skipping_do_not_change_a_bit(fp, n, k, kind, monkeypatch):
    """From 4 096 cells on the tensor-core path visits the cells along a Z-order curve and skips the
    products of weight digit planes that are all zero in a (cell tile x stage) block.  Sums of
    integers in another order, minus terms that are exactly zero: the output is BIT-IDENTICAL to
    the run in the given cell order, column for column; with FP_OUT_ANY_CELL_ORDER the same
    values come back in the library's order (same multiset per row, same median)."""
    import torch
    from scipy.stats import rankdata
    from fastproject_b200 import _native
    rng = np.random.RandomState(n + k)
    xy = np.concatenate([rng.normal(c, s, size=(n // 4, 2)) for c, s in ((-2, .3), (0, .5), (1.5, .2), (3, .4))])
    xy = np.concatenate([xy, rng.uniform(-4, 4, size=(n - xy.shape[0], 2))])
    coords = torch.from_numpy(fp.synth.unit_radius(xy)).cuda()
    ld = fp.E.padded(n)
    vals = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    if kind == 0:
        vals[:, :n] = torch.from_numpy(np.array([rankdata(rng.normal(size=n)) for _ in range(k)])).cuda()
    else:
        vals[:, :n] = torch.from_numpy((rng.uniform(size=(k, n)) < 0.3).astype(np.float64)).cuda()
    z = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="tc", output_kind=kind)[:, :n].cpu().numpy()
    free = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="tc", output_kind=kind,
                            any_cell_order=True)
    med_free = fp.E.row_medians(free, n).cpu().numpy()
    free = free[:, :n].cpu().numpy()
    monkeypatch.setenv("FASTPROJECT_B200_CELL_ORDER", "given")
    given = fp.E.consistency(coords, 0.33 ** 2, vals, n, precision="tc", output_kind=kind)[:, :n].cpu().numpy()
    assert not np.isnan(z).any()
    assert np.array_equal(z, given)
    assert np.array_equal(np.sort(free, axis=1), np.sort(given, axis=1))
    assert not np.array_equal(free, given)                       # it really is another order
    assert np.array_equal(med_free, np.median(given, axis=1))


def test_tensor_core_kernel_binary_predictions(fp):
    a, b = _tc_case(fp, 2000, 300, 5, binary=True, kind=1)
    assert np.abs(a - b).max() <= 1e-8          # values in [0, 1]: the scale puts +-1 in the top digit


def test_tensor_core_kernel_rejects_what_it_cannot_represent(fp):
    import torch
    from fastproject_b200 import _native
    n, k = 500, 3
    ld = fp.E.padded(n)
    proj, _ = fp.synth.make_projections(n, 1, seed=3)
    c = torch.from_numpy(proj["Proj00"]).cuda()
    vals = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    vals[:, :n] = torch.rand((k, n), dtype=torch.float64, device="cuda")      # not half-integers
    out = fp.E.consistency(c, 0.33 ** 2, vals, n, precision="tc")
    assert torch.isnan(out[:, :n]).all()                                      # loud, not silently wrong
    # a value range wider than three base-256 digits is refused the same way
    wide = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    wide[:, :n] = torch.arange(n, dtype=torch.float64, device="cuda") * 20000.0
    out = fp.E.consistency(c, 0.33 ** 2, wide, n, precision="tc")
    assert torch.isnan(out[:, :n]).all()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_sigs_vs_projections_midsize_vs_oracle(fp, precision):
    n, g = 3001, 800
    proj, grad = fp.synth.make_projections(n, 3, seed=21)
    x, w, z, genes, cells = fp.synth.make_expression(g, n, seed=22, planted=grad)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    defs = fp.synth.make_signature_dicts(genes, 40, seed=23)
    sigs = [fp.S.Signature(d, sg, "t", nm) for nm, d, sg in defs]
    random.seed(8); np.random.seed(8)
    rnd = fp.S.generate_random_sigs(genes, False, [5, 10, 20, 50, 100, 200], 50)
    real = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", z, 5)
    bg = fp.S.calculate_sig_scores(data, rnd, "weighted_avg", z, 5)
    got = fp.S.sigs_vs_projections(proj, real, bg, precision=precision)
    odata = O.ExpressionMatrix(x, w, genes, cells)
    oreal = O.calculate_sig_scores(odata, [O.Signature(s.sig_dict, s.signed, "t", s.name) for s in sigs],
                                   "weighted_avg", z, 5)
    obg = O.calculate_sig_scores(odata, [O.Signature(s.sig_dict, s.signed, "t", s.name) for s in rnd],
                                 "weighted_avg", z, 5)
    want = O.sigs_vs_projections(proj, oreal, obg)
    assert got[0] == want[0] and got[1] == want[1] and got[5] == want[5]
    np.testing.assert_allclose(got[2], want[2], rtol=_rtol(precision))
    np.testing.assert_allclose(got[4], want[4], rtol=_rtol(precision))
    np.testing.assert_allclose(got[3], want[3], rtol=LOGQ_RTOL, atol=_logq_atol(precision))
    assert np.array_equal(np.argsort(got[3], axis=None, kind="stable"),
                          np.argsort(want[3], axis=None, kind="stable"))
    assert np.array_equal(got[3] < -1.3, want[3] < -1.3)
    assert (want[3] < -1.3).any() and not (want[3] < -1.3).all()


def test_smoke_entry_point():
    import __graft_entry__ as ge
    with contextlib.redirect_stdout(io.StringIO()):
        ge.smoke()


def test_distributed_mode_single_rank_equals_local(fp):
    """distributed=True on a one-rank NCCL group takes the all-gather path and must return
    exactly what the local path returns (the two-rank exchange itself is covered on CPU by
    tests/test_distributed_cpu.py and on the GPUs by bench.py --gpus N)."""
    import os
    import torch.distributed as dist
    c = G.load("case_synth_even")
    want = _svp(fp, c, precision="fp64")
    created = False
    if not dist.is_initialized():
        import socket
        sock = socket.socket()
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
        sock.close()
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
        created = True
    try:
        got = _svp(fp, c, precision="fp64", distributed=True)
    finally:
        if created:
            dist.destroy_process_group()
    assert got[0] == want[0] and got[1] == want[1] and got[5] == want[5]
    for a, b in zip(got[2:5], want[2:5]):
        assert np.array_equal(a, b)


def test_full_size_properties_of_the_tensor_core_path(fp):
    """BASELINE config 2 cell count (20 000 cells): properties that do not need the O(N^2)
    reference.  (1) rows are independent: a duplicated rank row gives identical output;
    (2) the contraction is exact integer arithmetic, so permuting the cells (coordinates and
    rank columns together) permutes the output BIT FOR BIT; (3) a constant row is predicted
    to the precision of the digit classes; (4) medians agree with the FP64 verifier kernel on the same rows."""
    import torch
    from scipy.stats import rankdata
    n, k = 20000, 12
    rng = np.random.RandomState(77)
    proj, _ = fp.synth.make_projections(n, 1, seed=78)
    coords = proj["Proj00"]
    vals = np.array([rankdata(rng.normal(size=n) + (coords[0] if r % 2 else 0)) for r in range(k)])
    vals[5] = vals[2]                       # duplicate row
    vals[7] = 1234.5                        # constant row
    ld = fp.E.padded(n)

    def run(c, v, precision):
        dev = torch.zeros((v.shape[0], ld), dtype=torch.float64, device="cuda")
        dev[:, :n] = torch.from_numpy(v).cuda()
        out = fp.E.consistency(torch.from_numpy(np.ascontiguousarray(c)).cuda(), 0.33 ** 2, dev, n,
                               precision=precision)
        return out[:, :n].cpu().numpy()

    base = run(coords, vals, "tc")
    assert np.array_equal(base[5], base[2])
    # |v - prediction| of a constant row: zero up to the uncomputed lowest digit class (2^-32 relative)
    assert np.abs(base[7]).max() <= 1e-7 * n
    perm = rng.permutation(n)
    shuffled = run(coords[:, perm], vals[:, perm], "tc")
    assert np.array_equal(shuffled, base[:, perm])
    ref = run(coords, vals[:4], "fp64")
    np.testing.assert_allclose(np.median(base[:4], axis=1), np.median(ref, axis=1), rtol=0, atol=2e-4)
    assert np.abs(base[:4] - ref).max() <= 1e-7 * n


# ------------------------------------------------------------------ normalisation (next row N1)
@pytest.mark.parametrize("tag", ["small", "wide"])
def test_normalization_methods_bit_exact_vs_reference_golden(fp, tag):
    from fastproject_b200 import NormalizationMethods as NM
    c = G.load("case_norm")
    x = c[tag + "_X"]
    assert NM.no_normalization(x) is x
    assert np.array_equal(NM.row_normalization(x.copy()), c[tag + "_rows"])
    assert np.array_equal(NM.col_normalization(x.copy()), c[tag + "_cols"])
    assert np.array_equal(NM.row_and_col_normalization(x.copy()), c[tag + "_rows_cols"])
    assert np.array_equal(NM.col_rank_normalization(x.copy()), c[tag + "_col_rank"])


def test_normalization_midsize_bit_exact_vs_oracle_and_device_path(fp):
    """20 000 cells per gene (the pairwise tree is 8 levels deep), F-ordered input, and the
    device-resident variant feeding calculate_sig_scores without a host round trip."""
    import torch
    from fastproject_b200 import NormalizationMethods as NM
    rng = np.random.RandomState(5)
    x = np.log1p(rng.gamma(0.5, 4.0, size=(64, 20000))) * (rng.uniform(size=(64, 20000)) < 0.6)
    x[7] = 0.0
    assert np.array_equal(NM.row_normalization(x), O.row_normalization(x))
    assert np.array_equal(NM.col_normalization(x), O.col_normalization(x))
    assert np.array_equal(NM.row_normalization(np.asfortranarray(x)), O.row_normalization(x))
    small = x[:, :257]
    assert np.array_equal(NM.col_rank_normalization(small), O.col_rank_normalization(small))
    dev = NM.row_normalization(torch.from_numpy(x).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), O.row_normalization(x))
    genes = ["g%d" % i for i in range(64)]
    cells = ["c%d" % i for i in range(20000)]
    sig = fp.S.Signature(dict((g, 1) for g in genes[:20]), False, "t", "s")
    w = np.ones_like(x)
    got = fp.S.calculate_sig_scores(fp.synth.ExpressionMatrix(dev, torch.from_numpy(w).cuda(), genes, cells),
                                    [sig] * 1, "weighted_avg", None, 5)
    want = O.calculate_sig_scores(O.ExpressionMatrix(O.row_normalization(x), w, genes, cells),
                                  [O.Signature(sig.sig_dict, False, "t", "s")], "weighted_avg", None, 5)
    assert np.array_equal(got["s"].scores, want["s"].scores)


# ------------------------------------------------------------------ false-negative weights (next row N3)
WEIGHT_TOL = 1e-12      # absolute, weights in [0, 1]: exp() of CUDA and of NumPy differ in the last bit


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_compute_weights_vs_reference_golden(fp, tag):
    from fastproject_b200 import Transforms as T
    c = G.load("case_weights")
    x, params, want = c[tag + "_X"], c[tag + "_params"], c[tag + "_weights"]
    got = T.compute_weights(T.logistic_fnr, params, x)                 # p_nd = zero mask on the device
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=0, atol=WEIGHT_TOL, equal_nan=True)
    assert (got[x != 0] == 1.0).all()                                  # detected values: exactly 1
    # explicit p_nd (array and DataFrame selected by the data's labels, like the reference)
    import pandas as pd
    got2 = T.compute_weights(None, params, x, T.estimate_non_detection(x))
    assert np.array_equal(got2, got, equal_nan=True)
    genes = ["G%03d" % i for i in range(x.shape[0])]
    cells = ["c%04d" % i for i in range(x.shape[1])]
    df = T.estimate_non_detection(pd.DataFrame(x, index=genes, columns=cells))
    data = fp.synth.ExpressionMatrix(x, x * 0 + 1, genes, cells)
    shuffled = df.iloc[::-1, ::-1]                                     # label-based selection restores the order
    got3 = T.compute_weights(lambda v, x0, a, L=0, S=1: L + S / (1 + np.exp((v - x0) * a)), params, data, shuffled)
    assert np.array_equal(got3, got, equal_nan=True)
    with pytest.raises(NotImplementedError):
        T.compute_weights(lambda v, x0, a, L=0, S=1: np.tanh(v), params, x)


@pytest.mark.gpu
@pytest.mark.parametrize("g,n", [(5, 1), (7, 7), (9, 8), (33, 129), (40, 1001), (64, 4097), (300, 20000), (37, 19999), (6, 40001)])
def test_row_kernels_staged_in_shared_memory_equal_the_global_memory_ones(fp, g, n, monkeypatch):
    """Rows that fit in shared memory are staged there once (bulk asynchronous copies) by
    fp_normalize_rows and fp_compute_weights; longer rows, rows on odd pitches and an explicit
    p_nd take the kernels that read global memory.  Same tree, same operations: BIT-IDENTICAL,
    and both equal the oracle (bit for bit for the z-scores, 1e-12 for the weights)."""
    import torch
    from fastproject_b200 import _engine as E
    rng = np.random.RandomState(100 + n)
    x = np.log1p(rng.gamma(0.6, 5.0, size=(g, n))) * (rng.uniform(size=(g, n)) < 0.5)
    x[0] = 0.0                                            # never detected: NaN weights, sigma -> 1
    if g > 2:
        x[1] = 3.25                                       # always detected, constant
    params = np.vstack([rng.uniform(0.5, 4.0, size=n), rng.uniform(0.0, 2.0, size=n), np.zeros(n), np.ones(n)])
    xd = E.alloc_matrix(g, n, "cuda")                      # even pitch also when n is odd
    xd.copy_(torch.from_numpy(x))
    pd_ = torch.from_numpy(params).cuda()
    z_staged = E.normalize_rows(xd).cpu().numpy()[:, :n]
    w_staged = E.compute_weights(xd, pd_).cpu().numpy()[:, :n]
    monkeypatch.setenv("FASTPROJECT_B200_ROW_STAGE", "off")
    z_global = E.normalize_rows(xd).cpu().numpy()[:, :n]
    w_global = E.compute_weights(xd, pd_).cpu().numpy()[:, :n]
    monkeypatch.delenv("FASTPROJECT_B200_ROW_STAGE")
    assert np.array_equal(z_staged, z_global)
    assert np.array_equal(w_staged, w_global, equal_nan=True)
    assert np.array_equal(z_staged, O.row_normalization(x))
    with np.errstate(all="ignore"):
        want = O.compute_weights(O.logistic_fnr, params, x, O.estimate_non_detection(x))
    assert np.array_equal(np.isnan(w_staged), np.isnan(want))
    np.testing.assert_allclose(w_staged, want, rtol=0, atol=WEIGHT_TOL, equal_nan=True)
    # an odd pitch (rows not 16-byte aligned) falls back to the global-memory kernels by itself
    odd = torch.zeros((g, n + 1 + n % 2), dtype=torch.float64, device="cuda")[:, :n]
    odd.copy_(xd)
    assert odd.stride(0) % 2 == 1
    assert np.array_equal(E.normalize_rows(odd).cpu().numpy()[:, :n], z_staged)
    assert np.array_equal(E.compute_weights(odd, pd_).cpu().numpy()[:, :n], w_staged, equal_nan=True)


def test_compute_weights_midsize_vs_oracle_and_feeds_scoring(fp):
    """20 000 cells per gene, soft p_nd, device-resident output feeding calculate_sig_scores."""
    import torch
    from fastproject_b200 import Transforms as T
    rng = np.random.RandomState(9)
    g, n = 48, 20000
    x = np.log1p(rng.gamma(0.6, 5.0, size=(g, n))) * (rng.uniform(size=(g, n)) < 0.5)
    params = np.vstack([rng.uniform(0.5, 4.0, size=n), rng.uniform(0.0, 2.0, size=n), np.zeros(n), np.ones(n)])
    want = O.compute_weights(O.logistic_fnr, params, x, O.estimate_non_detection(x))
    got = T.compute_weights(T.logistic_fnr, params, x)
    np.testing.assert_allclose(got, want, rtol=0, atol=WEIGHT_TOL)
    soft = np.clip(O.estimate_non_detection(x) * 0.9 + 0.05 * rng.uniform(size=(g, n)), 0, 1)
    want = O.compute_weights(O.logistic_fnr, params, x, soft)
    np.testing.assert_allclose(T.compute_weights(None, params, x, soft), want, rtol=0, atol=WEIGHT_TOL)
    w_dev = T.compute_weights(None, torch.from_numpy(params).cuda(), torch.from_numpy(x).cuda())
    assert w_dev.is_cuda
    np.testing.assert_allclose(w_dev.cpu().numpy(), O.compute_weights(
        O.logistic_fnr, params, x, O.estimate_non_detection(x)), rtol=0, atol=WEIGHT_TOL)
    genes = ["g%d" % i for i in range(g)]
    cells = ["c%d" % i for i in range(n)]
    sig = fp.S.Signature(dict((q, 1) for q in genes[:20]), False, "t", "s")
    res = fp.S.calculate_sig_scores(fp.synth.ExpressionMatrix(torch.from_numpy(x).cuda(), w_dev, genes, cells),
                                    [sig], "weighted_avg", None, 5)
    ref = O.calculate_sig_scores(O.ExpressionMatrix(x, w_dev.cpu().numpy(), genes, cells),
                                 [O.Signature(sig.sig_dict, False, "t", "s")], "weighted_avg", None, 5)
    assert np.array_equal(res["s"].scores, ref["s"].scores)


# ------------------------------------------------------------------ Interactive wrappers (next row N4)
def test_interactive_dataframe_wrappers_vs_oracle(fp):
    import pandas as pd
    from fastproject_b200 import Interactive as I
    c = G.load("case_synth_even")
    genes = [str(g) for g in c["genes"]]
    cells = [str(x) for x in c["cells"]]
    data = pd.DataFrame(c["X"], index=genes, columns=cells)
    weights = pd.DataFrame(c["W"], index=genes, columns=cells)
    real = G.unpack_sigs(c, "real_", fp.S.Signature)
    rnd = G.unpack_sigs(c, "rnd_", fp.S.Signature)
    scores, num_genes = I.calculate_sig_scores(data, real, weights, "weighted_avg", c["Z"], int(c["min_genes"]))
    rscores, rnum = I.calculate_sig_scores(data, rnd, weights, "weighted_avg", c["Z"], int(c["min_genes"]))
    odata = O.ExpressionMatrix(c["X"], c["W"], genes, cells)
    oreal = O.calculate_sig_scores(odata, G.unpack_sigs(c, "real_", O.Signature), "weighted_avg", c["Z"],
                                   int(c["min_genes"]))
    ornd = O.calculate_sig_scores(odata, G.unpack_sigs(c, "rnd_", O.Signature), "weighted_avg", c["Z"],
                                  int(c["min_genes"]))
    assert list(scores.columns) == list(oreal.keys()) and list(scores.index) == cells
    for name in oreal:
        assert np.array_equal(scores[name].values, oreal[name].scores)
        assert num_genes[name] == oreal[name].numGenes
    proj_name = sorted(G.projections(c).keys())[0]
    coords = G.projections(c)[proj_name]
    projection = pd.DataFrame(coords.T, index=cells, columns=["x", "y"])
    cons, pvals, null = I.sigs_vs_projection(projection, scores, num_genes, rscores, rnum)
    want = O.sigs_vs_projections({"proj_name": coords}, oreal, ornd)
    assert list(cons.index) == want[0] and list(null.index) == want[5]
    np.testing.assert_allclose(cons.values, want[2], rtol=TC_CONS_RTOL)
    np.testing.assert_allclose(null.values, want[4], rtol=TC_CONS_RTOL)
    np.testing.assert_allclose(pvals.values, want[3], rtol=LOGQ_RTOL, atol=1e-6)


def test_prefetch_scores_gives_the_same_host_arrays(fp):
    """Signatures.prefetch_scores starts the device -> host copy of a result on a side stream;
    the arrays read afterwards are those of the synchronous path, and the oracle's."""
    x, w, z, genes, cells = fp.synth.make_expression(300, 257, seed=31)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    sigs = [fp.S.Signature(d, sg, "t", nm) for nm, d, sg in fp.synth.make_signature_dicts(genes, 12, seed=4)]
    plain = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", z, 2)
    ahead = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", z, 2)
    fp.S.prefetch_scores(ahead)
    fp.S.prefetch_scores(ahead)                                   # idempotent
    want = O.calculate_sig_scores(O.ExpressionMatrix(x, w, genes, cells),
                                  [O.Signature(s.sig_dict, s.signed, "t", s.name) for s in sigs], "weighted_avg", z, 2)
    assert list(plain) == list(ahead) == list(want)
    for name in want:
        assert np.array_equal(ahead[name].scores, plain[name].scores)
        assert np.array_equal(ahead[name].scores, want[name].scores)
        assert np.array_equal(ahead[name].ranks, plain[name].ranks)
    fp.S.prefetch_scores(dict())


@pytest.mark.parametrize("n,k,kw", [(1, 2, {}), (129, 127, {}), (1001, 130, dict(outliers=1)), (4097, 260, dict(dup=3)),
                                    (20000, 9, {}), (1000, 6, dict(binary=True))])
def test_value_planes_staged_in_shared_memory_do_not_change_a_bit(fp, n, k, kw, monkeypatch):
    """The value digit planes of the tensor-core path are cut from rows staged in shared memory
    (rows that fit) or gathered from global memory: same planes, so the outputs are identical."""
    a, staged = _tc_case(fp, n, k, 300 + n, **kw)
    monkeypatch.setenv("FASTPROJECT_B200_ROW_STAGE", "off")
    _, gathered = _tc_case(fp, n, k, 300 + n, **kw)
    assert np.array_equal(staged, gathered)
    assert np.abs(a - staged).max() <= 1e-7 * max(n, 16)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_single_dissimilarity_buffer_gives_the_same_result(fp, precision, monkeypatch):
    """Above MedianOverlap.DOUBLE_BUFFER_MAX bytes (a million cells) the medians and the next
    contraction share ONE dissimilarity buffer and take turns; forcing that at test size must
    not change a bit of the result."""
    c = G.load("case_synth_odd")
    two = _svp(fp, c, precision=precision)
    monkeypatch.setattr(fp.E.MedianOverlap, "DOUBLE_BUFFER_MAX", 0)
    one = _svp(fp, c, precision=precision)
    assert two[0] == one[0] and two[1] == one[1] and two[5] == one[5]
    for a, b in zip(two[2:5], one[2:5]):
        assert np.array_equal(a, b)
    _check_svp(one, c, "svp", rtol=_rtol(precision))


@pytest.mark.parametrize("n,k,kw", [(1, 2, {}), (129, 127, {}), (1000, 5, dict(outliers=2)), (4097, 260, dict(dup=3)),
                                    (6000, 40, dict(per_cell=True)), (20000, 9, {}), (33000, 3, dict(outliers=1)),
                                    (5000, 6, dict(binary=True))])
def test_packed_values_give_the_same_bits_as_the_matrix(fp, n, k, kw):
    """fp_values_pack + fp_consistency_packed (the value range and the digit planes made once,
    permuted per projection) against fp_consistency on the matrix itself: identical outputs, in
    the caller's cell order, in any cell order (as multisets per row) and for the raw prediction;
    two projections from one packed buffer."""
    import torch
    from scipy.stats import rankdata
    rng = np.random.RandomState(500 + n)
    proj, _ = fp.synth.make_projections(n, 2, seed=n)
    ld = fp.E.padded(n)
    if kw.get("binary"):
        vals = (rng.uniform(size=(k, n)) < 0.3).astype(np.float64)
    else:
        vals = np.array([rankdata(np.round(rng.normal(size=n), 2 if r % 2 else 6)) for r in range(k)])
    dev = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(vals).cuda()
    packed = fp.E.pack_values(dev, n, precision="tc")
    assert packed is not None and (packed.k, packed.n) == (k, n)
    for name in sorted(proj):
        c = torch.from_numpy(proj[name].copy()).cuda()
        if kw.get("outliers"):
            c[:, 0] = torch.tensor([7.0, -6.0], dtype=torch.float64)
        if kw.get("dup"):
            c[:, n - 1] = c[:, 5]
        sig = 0.33 ** 2
        if kw.get("per_cell"):
            sig = torch.from_numpy((0.2 + 0.3 * rng.uniform(size=n)) ** 2).cuda()
        for kind in (0, 1):                                   # |rank - prediction|, prediction
            a = fp.E.consistency(c, sig, dev, n, precision="tc", output_kind=kind)[:, :n].cpu().numpy()
            b = fp.E.consistency(c, sig, dev, n, precision="tc", output_kind=kind, packed=packed)[:, :n].cpu().numpy()
            assert np.array_equal(a, b)
        a = fp.E.consistency(c, sig, dev, n, precision="tc", any_cell_order=True)[:, :n].cpu().numpy()
        b = fp.E.consistency(c, sig, dev, n, precision="tc", any_cell_order=True, packed=packed)[:, :n].cpu().numpy()
        assert np.array_equal(a, b)
        assert not fp.E.consistency_refused(dev, n, precision="tc")
    with pytest.raises(ValueError):
        fp.E.consistency(c, 0.1, dev[:1], n, precision="tc", packed=packed)
    # FP64 mode ignores the packed form (and nothing is packed for it)
    assert fp.E.pack_values(dev, n, precision="fp64") is None


def test_packed_values_carry_the_refusal(fp):
    """Values the tensor-core mode cannot hold exactly are refused at packing time; every call on
    the packed buffer then returns NaN and reports it, like the call on the matrix."""
    import torch
    n, k = 5000, 4
    rng = np.random.RandomState(3)
    proj, _ = fp.synth.make_projections(n, 1, seed=2)
    c = torch.from_numpy(proj["Proj00"]).cuda()
    vals = rng.normal(size=(k, n))                            # not half-integers
    dev = torch.zeros((k, fp.E.padded(n)), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(vals).cuda()
    packed = fp.E.pack_values(dev, n, precision="tc")
    out = fp.E.consistency(c, 0.33 ** 2, dev, n, precision="tc", packed=packed, slot="refusal")[:, :n]
    assert fp.E.consistency_refused(dev, n, precision="tc", slot="refusal")
    assert bool(torch.isnan(out).all())


def test_integration_md_ctypes_stub_runs(fp):
    """The ctypes binding printed in INTEGRATION.md section 2 is executed as it stands (only the
    library path is made absolute): its scores equal the mirror's, its medians -- one projection
    at a time and through the packed-values loop -- equal each other and the engine's."""
    import os
    import re
    import torch
    from scipy.stats import rankdata
    from fastproject_b200 import _native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# FastProject/_b200\.py.*?)```", text, re.S).group(1)
    block = block.replace('ctypes.CDLL("libfastproject_b200.so")', "ctypes.CDLL(%r)" % _native.LIB_PATH)
    ns = dict()
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    n, k = 5000, 40
    rng = np.random.RandomState(8)
    proj, _ = fp.synth.make_projections(n, 2, seed=3)
    coords = [torch.from_numpy(proj[name].copy()).cuda() for name in sorted(proj)]
    ranks = torch.from_numpy(np.array([rankdata(rng.normal(size=n)) for _ in range(k)])).cuda()
    meds = ns["median_dissimilarities"](coords, ranks)
    for i, c in enumerate(coords):
        one = ns["median_dissimilarity"](c, ranks)
        assert torch.equal(meds[i], one)
        dis = fp.E.consistency(c, 0.33 ** 2, ranks, n, precision="tc")
        assert torch.equal(one, fp.E.row_medians(dis, n))
    x, w, z, genes, cells = fp.synth.make_expression(300, 256, seed=5)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    sigs = [fp.S.Signature(d, sg, "t", nm) for nm, d, sg in fp.synth.make_signature_dicts(genes, 6, seed=2)]
    from fastproject_b200 import _scoring
    kept, ptr_, gene, sign, _ = _scoring.pack_signatures(sigs, genes, 1)
    got = ns["weighted_scores"](data, ptr_, gene, sign).cpu().numpy()
    want = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", z, 1)
    assert np.array_equal(got, np.array([want[sigs[p].name].scores for p in kept]))
