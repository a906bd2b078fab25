"""Host-side logic of the mirror (CSR packing, grouping, BH, RNG streams) against
the oracle and the golden vectors.  CPU only."""
import random

import numpy as np
import pytest

from fastproject_b200 import Signatures as S
from fastproject_b200 import _scoring


def test_packer_extension_long_lists_odd_values_and_fallbacks():
    """csrc/pack_ext.c against the per-signature path: lists long enough for the bitmap ordering
    (> 32 genes), signed entries, float / bool / numpy values, ignored values, names that are not
    on the matrix, rejected signatures, an empty list; inputs it declines (numpy integers as
    values, duplicated labels, dict subclasses) must come back as None, not as a wrong table."""
    from fastproject_b200 import _native
    ext = _native.pack_ext()

    class Sig(object):
        def __init__(self, d):
            self.sig_dict = d
    rng = np.random.RandomState(11)
    genes = ["gene_%d" % i for i in range(5000)]
    values = [1, -1, 0, 1.0, -1.0, 0.0, True, 2, 0.5, np.float64(-1.0), -7, 10 ** 30]
    sigs = []
    for size in (1, 2, 5, 31, 32, 33, 48, 49, 100, 640, 3000):
        d = dict((genes[j], values[int(rng.randint(len(values)))]) for j in rng.choice(5000, size, replace=False))
        d["not_on_the_matrix_%d" % size] = 1
        sigs.append(Sig(d))
    sigs.append(Sig({}))
    sigs.append(Sig({"nowhere": 1}))
    index = _scoring.gene_row_index(genes)
    table = ext.label_table(genes)
    for min_genes in (0, 3, 40, 1e99):
        fast = _scoring.pack_signatures(sigs, genes, min_genes)
        slow = _scoring._pack_signatures_slow(sigs, index, min_genes)
        assert fast[0] == slow[0]
        for a, b in zip(fast[1:], slow[1:]):
            assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)
        assert ext.pack([s.sig_dict for s in sigs], table, float(min_genes)) is not None
    assert [len(np.frombuffer(b, dtype=np.uint8)) for b in ext.pack([], table, 0.0)] == [0, 4, 0, 0, 0]
    assert ext.pack([{genes[0]: np.int64(1)}], table, 0.0) is None
    assert ext.pack([{genes[0]: "1"}], table, 0.0) is None
    assert ext.pack([{genes[0]: 1, 17: 1}], table, 0.0) is None         # a name that is not a str
    assert ext.label_table(genes[:10] + genes[:10]) is None             # duplicated labels: no table
    assert ext.label_table(genes[:10] + [3]) is None                    # labels that are not str
    assert ext.label_table(tuple(genes)) is None
    with pytest.raises(ValueError):
        ext.pack([{genes[0]: 1}], index, 0.0)                           # not a table
    import collections
    assert ext.pack([collections.OrderedDict({genes[0]: 1})], table, 0.0) is None
    # equal names held by OTHER str objects (signatures read from a file) match by content
    copies = dict(("gene_" + str(i), 1) for i in (5, 17, 4999))
    got = ext.pack([copies], table, 0.0)
    assert list(np.frombuffer(got[2], dtype=np.int32)) == [5, 17, 4999]
    # the cached table follows the CONTENT of the label list, not the object
    labels = list(genes)
    a = _scoring.pack_signatures(sigs, labels, 0)
    labels[0], labels[1] = labels[1], labels[0]
    b = _scoring.pack_signatures(sigs, labels, 0)
    want = _scoring._pack_signatures_slow(sigs, _scoring.gene_row_index(labels), 0)
    assert all(np.array_equal(x, y) for x, y in zip(b[1:], want[1:]))
    assert not all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
    # ... and the public packer still answers for all of them
    odd = [Sig({genes[0]: np.int64(1), genes[7]: np.int64(-1)}), Sig(collections.OrderedDict({genes[3]: 1}))] * 5
    fast = _scoring.pack_signatures(odd, genes, 0)
    slow = _scoring._pack_signatures_slow(odd, index, 0)
    assert fast[0] == slow[0] and all(np.array_equal(a, b) for a, b in zip(fast[1:], slow[1:]))


from oracle import fp_oracle as O
from tests import _golden as G


@pytest.mark.parametrize("case_name", ["case_synth_odd", "case_synth_even", "case_fixture"])
def test_csr_packing_equals_sig_indices(case_name):
    c = G.load(case_name)
    genes = [str(g) for g in c["genes"]]
    sigs = G.unpack_sigs(c, "real_", S.Signature) + G.unpack_sigs(c, "rnd_", S.Signature)
    kept, ptr, gene, sign, counts = _scoring.pack_signatures(sigs, genes, 0)
    osigs = [O.Signature(s.sig_dict, s.signed, s.source, s.name) for s in sigs]
    row = 0
    for pos, (s, o) in enumerate(zip(sigs, osigs)):
        dense = o.sig_indices(genes)
        assert np.array_equal(s.sig_indices(genes), dense)
        nz = np.nonzero(dense)[0]
        if nz.size == 0:
            assert pos not in kept
            continue
        assert kept[row] == pos
        assert np.array_equal(gene[ptr[row]:ptr[row + 1]], nz)
        assert np.array_equal(sign[ptr[row]:ptr[row + 1]], dense[nz, 0])
        assert counts[row] == nz.size
        row += 1
    assert row == len(kept)


def test_min_genes_filter_and_errors():
    genes = ["a", "b", "c", "d"]
    sigs = [S.Signature({"a": 1, "b": -1, "c": 0}, True, "f", "three"),
            S.Signature({"a": 1}, False, "f", "one"),
            S.Signature({"zz": 1}, False, "f", "none"),
            S.Signature({"a": 0.5, "b": 2}, False, "f", "ignored_values")]
    kept, ptr, gene, sign, counts = _scoring.pack_signatures(sigs, genes, 2)
    assert kept == [0] and list(counts) == [3] and list(sign) == [1, -1, 1]
    with pytest.raises(ValueError, match="Too few genes"):
        _scoring.pack_signatures([sigs[1]], genes, 2, raise_on_reject=True)
    with pytest.raises(ValueError, match="No genes match"):
        _scoring.pack_signatures([sigs[2]], genes, 0, raise_on_reject=True)
    with pytest.raises(ValueError, match="No genes match"):
        _scoring.pack_signatures([sigs[3]], genes, 0, raise_on_reject=True)
    # the function default of calculate_sig_scores (1e99) drops everything
    assert _scoring.pack_signatures(sigs, genes, 1e99)[0] == []


def test_duplicate_gene_labels_all_match():
    genes = ["a", "b", "a", "c"]
    s = S.Signature({"a": -1, "c": 1}, True, "f", "dup")
    o = O.Signature(s.sig_dict, True, "f", "dup")
    assert np.array_equal(s.sig_indices(genes), o.sig_indices(genes))


def test_background_groups_match_oracle_lookup():
    rng = np.random.RandomState(3)
    rnd_ng = [int(v) for v in rng.choice([5, 10, 20, 50], size=60)]
    real_ng = [1, 5, 7, 8, 15, 35, 49, 51, 1000]
    order, ptr, grp = S._background_groups(rnd_ng, real_ng)
    med = rng.uniform(size=60)
    table = O.background_table(med, rnd_ng)
    assert len(ptr) - 1 == table.shape[0]
    for g in range(table.shape[0]):
        members = order[ptr[g]:ptr[g + 1]]
        assert all(rnd_ng[m] == table[g, 0] for m in members)
        assert list(members) == sorted(members)
        assert np.mean(med[members]) == table[g, 1]
    for k, ng in enumerate(real_ng):
        assert grp[k] == int(np.argmin(np.abs(ng - table[:, 0])))
    with pytest.raises(ValueError):
        S._background_groups([], [5])


def test_p_to_q_golden():
    c = G.load("case_randsigs")
    assert np.array_equal(S.p_to_q(c["p_in"].copy()), c["q_out"])


def test_generate_random_sigs_rng_streams():
    c = G.load("case_randsigs")
    genes = [str(g) for g in c["genes"]]
    random.seed(424242)
    np.random.seed(171717)
    a = S.generate_random_sigs(genes, False, None, 3)
    b = S.generate_random_sigs(genes, True, [7, 13], 5)
    for got, tag in ((a, "a_"), (b, "b_")):
        want = G.unpack_sigs(c, tag, O.Signature)
        assert [s.name for s in got] == [s.name for s in want]
        for s, t in zip(got, want):
            assert list(s.sig_dict.items()) == list(t.sig_dict.items())
            assert s.signed == t.signed and s.source == "x"


def test_import_seeds_global_random_like_reference():
    import importlib
    import fastproject_b200.Global as Gl
    importlib.reload(Gl)
    first = random.random()
    random.seed(1335607828)
    assert first == random.random()


def test_result_writers_byte_identical_to_reference(tmp_path):
    """fastproject_b200.FileIO against text written by the live reference's FileIO.write_matrix /
    write_signature_scores (tests/golden/make_golden.py writers_case)."""
    from fastproject_b200 import FileIO
    c = G.load("case_writers")
    data, rows, cols = c["data"], [str(r) for r in c["rows"]], [str(x) for x in c["cols"]]
    FileIO.write_matrix(str(tmp_path / "m.txt"), data, rows, cols)
    assert open(str(tmp_path / "m.txt")).read() == str(c["matrix_text"])

    class Scores(object):
        def __init__(self, scores, is_factor):
            self.scores, self.isFactor = scores, is_factor
    sigs = dict()
    for i, name in enumerate(str(n) for n in c["score_order"]):
        sigs[name] = Scores(["a", "b", "c", "a", "b"] * 5, True) if name == "F" else Scores(data[i], False)
    FileIO.write_signature_scores(str(tmp_path / "s.txt"), sigs, cols)
    assert open(str(tmp_path / "s.txt")).read() == str(c["scores_text"])
    FileIO.write_path_results(str(tmp_path / "out"), sigs, cols, data, data, rows, cols)
    assert open(str(tmp_path / "out" / "PMatrix.txt")).read() == str(c["matrix_text"])
    assert open(str(tmp_path / "out" / "SignatureScores.txt")).read() == str(c["scores_text"])


def test_significance_filter_semantics():
    """Pipelines.filter_significant_signatures (restatement of Pipelines.py:299-349; unpinned: the
    reference block does not run on this pandas).  Checks the rules the block states."""
    from fastproject_b200 import Pipelines as P

    class Scores(object):
        def __init__(self, pre):
            self.isPrecomputed = pre
    rng = np.random.RandomState(4)
    keys = ["sig%03d" % i for i in rng.permutation(300)]

    def model(scale):
        def proj():
            return dict(sigProjMatrix=rng.normal(size=(300, 4)),
                        sigProjMatrix_p=-np.abs(rng.normal(size=(300, 4))) * scale, signatureKeys=list(keys))
        return dict(signatureScores=dict((k, Scores(i % 50 == 0)) for i, k in enumerate(keys)),
                    projectionData=[proj(), proj()])
    # few samples, many significant: everything with min log10 q <= -1.3 stays (+ precomputed)
    m = model(3.0)
    names, vals = P.signature_significance(m["projectionData"])
    assert names == sorted(keys)
    want = set(k for k, v in zip(names, vals) if v <= -1.3) | set(k for k in keys if m["signatureScores"][k].isPrecomputed)
    keep = P.filter_significant_signatures(m, n_samples=500)
    assert set(k for k in keep if keep[k]) == want and len(want) > 200
    for pd_ in m["projectionData"]:
        assert pd_["signatureKeys"] == [k for k in keys if k in want]
        assert pd_["sigProjMatrix"].shape == (len(want), 4) == pd_["sigProjMatrix_p"].shape
    # few significant: the threshold rises to the 200th best value
    m = model(0.2)
    keep = P.filter_significant_signatures(m, n_samples=500)
    n_kept_scored = sum(1 for k in keep if keep[k] and not m["signatureScores"][k].isPrecomputed)
    assert 194 <= n_kept_scored <= 200
    # many samples: hard limit of 200; all_sigs keeps everything
    m = model(3.0)
    keep = P.filter_significant_signatures(m, n_samples=5000)
    assert sum(1 for k in keep if keep[k] and not m["signatureScores"][k].isPrecomputed) <= 200
    m = model(3.0)
    assert all(P.filter_significant_signatures(m, n_samples=5000, all_sigs=True).values())


def _grid_nearest_neighbour(x, y):
    """Line-by-line Python model of nn_count / nn_search_kernel (csrc/consistency_tc.cu): grid over
    mean +- 3 sigma clipped to the extremes, cells outside it in the border grid cells, Chebyshev
    rings until the best distance is within the block searched so far."""
    import math
    n = len(x)
    G = int(min(4096, math.ceil(math.sqrt(n))))

    def axis(v):
        mean, sd = v.mean(), v.std()
        a, b = max(v.min(), mean - 3 * sd), min(v.max(), mean + 3 * sd)
        if not b > a:
            a, b = v.min(), v.max()
        w = b - a
        return a, (w / G if 0 < w < 1e300 else 1.0)
    (xmin, hx), (ymin, hy) = axis(x), axis(y)

    def coord(v, vmin, h):
        t = (v - vmin) * (1.0 / h)
        return np.clip(np.where(t >= G, G - 1, np.trunc(t)), 0, G - 1).astype(np.int64)
    cx, cy = coord(x, xmin, hx), coord(y, ymin, hy)
    cell = cy * G + cx
    order = np.argsort(cell, kind="stable")
    start = np.concatenate([[0], np.cumsum(np.bincount(cell, minlength=G * G))])
    sx, sy = x[order], y[order]
    out = np.full(n, np.inf)
    for t in range(n):
        px, py, pcx, pcy = sx[t], sy[t], cx[order[t]], cy[order[t]]
        best, r = np.inf, 0
        while True:
            x0, x1, y0, y1 = pcx - r, pcx + r, pcy - r, pcy + r
            for uy in range(max(0, y0), min(G - 1, y1) + 1):
                if uy in (y0, y1):
                    segs = [(max(0, x0), min(G - 1, x1))]
                else:
                    segs = [(u, u) for u in ((x0,) if r == 0 else (x0, x1)) if 0 <= u <= G - 1]
                for ca, cb in segs:
                    for q in range(start[uy * G + ca], start[uy * G + cb + 1]):
                        if q != t:
                            dx, dy = px - sx[q], py - sy[q]
                            best = min(best, dx * dx + dy * dy)
            if best == 0.0:
                break
            bound = np.inf
            if x0 > 0:
                bound = min(bound, px - (xmin + x0 * hx))
            if x1 < G - 1:
                bound = min(bound, (xmin + (x1 + 1) * hx) - px)
            if y0 > 0:
                bound = min(bound, py - (ymin + y0 * hy))
            if y1 < G - 1:
                bound = min(bound, (ymin + (y1 + 1) * hy) - py)
            if not bound < np.inf:
                break
            bound = max(bound, 0.0) * (1.0 - 1e-9)
            if best <= bound * bound:
                break
            r += 1
        out[order[t]] = best
    return out


def test_nearest_neighbour_grid_search_rule():
    """The ring-search termination rule and the outlier-proof grid box of the device kernels,
    checked against the O(N^2) minimum on layouts that stress them (the kernels themselves are
    compared with the brute-force kernel in tests/test_gpu_parity.py)."""
    rng = np.random.RandomState(0)
    layouts = [
        (rng.uniform(-1, 1, 700), rng.uniform(-1, 1, 700)),
        (np.concatenate([rng.normal(c, 0.02, 150) for c in (-1, 0, 0.3, 2)]),
         np.concatenate([rng.normal(c, 0.03, 150) for c in (1, 0, -0.3, 2)])),
        (np.r_[rng.normal(0, 0.3, 500), 900.0, -700.0], np.r_[rng.normal(0, 0.3, 500), -500.0, 600.0]),
        (np.r_[rng.normal(0, 0.3, 300), np.zeros(50)], np.r_[rng.normal(0, 0.3, 300), np.zeros(50)]),
        (np.linspace(0, 1, 300), np.zeros(300)),
        (np.ones(40), np.ones(40)),
        (np.array([0.0, 1.0, 10.0]), np.array([0.0, 0.0, 0.0])),
    ]
    for x, y in layouts:
        x, y = np.asarray(x, float), np.asarray(y, float)
        d = (x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2
        np.fill_diagonal(d, np.inf)
        assert np.array_equal(_grid_nearest_neighbour(x, y), d.min(axis=1))


def test_fp_match_labels_host_function():
    """The library's host-side label matcher (Signature.sig_indices, Signatures.py:738-753)."""
    import __graft_entry__ as ge
    ge.build()
    from fastproject_b200 import _native, _scoring
    lib = _native.load()

    def match(labels, queries):
        lb, qb = "\0".join(labels).encode("utf-8"), "\0".join(queries).encode("utf-8")
        out = np.full(len(queries), -7, dtype=np.int32)
        rc = lib.fp_match_labels(lb, len(lb), len(labels), qb, len(qb), len(queries), out.ctypes.data)
        return rc, out

    labels = ["ACTB", "", "Gene-é", "A", "AB", "actb"]
    rc, rows = match(labels, ["AB", "ACTB", "missing", "", "Gene-é", "actb", "A", "ACT"])
    assert rc == 0 and rows.tolist() == [4, 0, -1, 1, 2, 5, 3, -1]
    rc, _ = match(["X", "Y", "X"], ["X"])
    assert rc == 1                                      # duplicated row label: caller's slow path
    rc, rows = match([], ["X", "Y"])
    assert rc == 0 and rows.tolist() == [-1, -1]
    lb = b"A\0B"
    out = np.zeros(1, dtype=np.int32)
    assert lib.fp_match_labels(lb, len(lb), 3, b"A", 1, 1, out.ctypes.data) == -1     # fewer labels than stated
    assert lib.fp_match_labels(lb, len(lb), 1, b"A", 1, 1, out.ctypes.data) == -1     # more labels than stated
    # the packer built on it: duplicated labels, odd values, non-string names all agree with the slow path
    class Sig(object):
        def __init__(self, d, name):
            self.sig_dict, self.name = d, name
    rng = np.random.RandomState(3)
    genes = ["g%d" % i for i in range(60)]
    sigs = [Sig(dict((genes[j], int(rng.choice([1, -1, 0]))) for j in rng.choice(60, 9, replace=False)), "s%d" % k)
            for k in range(12)]
    sigs[3].sig_dict["g5"] = 0.5                        # ignored value (weighted GMT entries)
    sigs[4].sig_dict["nope"] = 1
    for labels2 in (genes, genes[:30] + genes[:30], [str(g) for g in genes]):
        fast = _scoring.pack_signatures(sigs, labels2, 3)
        slow = _scoring._pack_signatures_slow(sigs, _scoring.gene_row_index(labels2), 3)
        for a, b in zip(fast, slow):
            assert np.array_equal(np.asarray(a), np.asarray(b))
    sigs[5].sig_dict[17] = 1                            # a non-string name: falls back, same answer
    fast = _scoring.pack_signatures(sigs, genes, 3)
    slow = _scoring._pack_signatures_slow(sigs, _scoring.gene_row_index(genes), 3)
    for a, b in zip(fast, slow):
        assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("case", ["case_filter_small", "case_filter_large"])
def test_significance_filter_vs_reference_golden(case):
    """Pipelines.Analysis' pruning of insignificant signatures (Pipelines.py:299-349).  The golden
    files hold what went into that block (every sigs_vs_projections output) and what came out of
    it in a run of the UNMODIFIED reference pipeline (tests/config1/run_config1.py golden, with the
    pandas / open() shims listed there): small = 260 samples, 262 signatures (threshold -1.3 raised
    to the 200th best value), large = more than 2000 samples (the 200 best signatures)."""
    import os
    from fastproject_b200 import Pipelines as P
    path = os.path.join(os.path.dirname(__file__), "golden", case + ".npz")
    if not os.path.exists(path):
        pytest.skip(case + " not generated")
    d = np.load(path)

    class Score(object):
        def __init__(self, pre):
            self.isPrecomputed = bool(pre)
    model = dict(signatureScores=dict((str(n), Score(p)) for n, p in zip(d["sig_names"], d["sig_precomputed"])),
                 projectionData=[])
    n_calls = int(d["n_calls"])
    for i in range(n_calls):
        model["projectionData"].append(dict(signatureKeys=[str(k) for k in d["before_keys_%d" % i]],
                                            sigProjMatrix=d["before_cons_%d" % i].copy(),
                                            sigProjMatrix_p=d["before_logq_%d" % i].copy()))
    keep = P.filter_significant_signatures(model, int(d["n_samples"]))
    assert sum(keep.values()) < len(keep)               # the case does prune
    for i in range(n_calls):
        got = model["projectionData"][i]
        assert got["signatureKeys"] == [str(k) for k in d["after_keys_%d" % i]]
        assert np.array_equal(got["sigProjMatrix"], d["after_cons_%d" % i])
        assert np.array_equal(got["sigProjMatrix_p"], d["after_logq_%d" % i])
