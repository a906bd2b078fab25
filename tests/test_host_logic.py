"""Host-side logic of the mirror (CSR packing, grouping, BH, RNG streams) against
the oracle and the golden vectors.  CPU only."""
import random

import numpy as np
import pytest

from fastproject_b200 import Signatures as S
from fastproject_b200 import _scoring
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
