"""The whole-job CPU driver of bench.py's CPU legs (oracle/cpu_job.py): the installed reference
(baseline/_ref, when present) and the row-blocked port must tell the same story, whole or dealt
to steps."""
import random

import numpy as np
import pytest

from oracle import cpu_job as J, reference as R
from fastproject_b200 import synth, Signatures as S


def _job_inputs(n=240, g=300):
    proj, grad = synth.make_projections(n, 2, seed=4)
    x, w, z, genes, cells = synth.make_expression(g, n, seed=5, planted=grad)
    defs = synth.make_signature_dicts(genes, 12, seed=6)
    real = [S.Signature(d, sg, "t", nm) for nm, d, sg in defs]
    random.seed(2); np.random.seed(2)
    rnd = S.generate_random_sigs(genes, False, [5, 10, 20], 15)
    return x, w, genes, cells, proj, real, rnd


def test_port_whole_equals_port_in_steps_and_oracle():
    from oracle import fp_oracle as O
    x, w, genes, cells, proj, real, rnd = _job_inputs()
    whole = J.CpuJob(x, w, genes, cells, proj, real, rnd, kind="port", chunk=7, row_block=64)
    assert whole.kind == "port"
    whole.run()
    parts = J.CpuJob(x, w, genes, cells, proj, real, rnd, kind="port", chunk=7, row_block=64)
    bounds = parts.step_bounds(5)
    assert bounds[0][0] == 0 and bounds[-1][1] == len(parts.units)
    for lo, hi in bounds:
        parts.run(lo, hi)
    for a, b in ((whole.cons, parts.cons), (whole.logq, parts.logq), (whole.rnd_cons, parts.rnd_cons)):
        assert np.array_equal(a, b)
    odata = O.ExpressionMatrix(x, w, genes, cells)
    conv = lambda sigs: [O.Signature(s.sig_dict, s.signed, s.source, s.name) for s in sigs]
    want = O.sigs_vs_projections(proj, O.calculate_sig_scores(odata, conv(real), "weighted_avg", None, 5),
                                 O.calculate_sig_scores(odata, conv(rnd), "weighted_avg", None, 5))
    assert whole.row_names == want[0] and whole.rnd_names == want[5]
    np.testing.assert_allclose(whole.cons, want[2], rtol=1e-12)
    np.testing.assert_allclose(whole.logq, want[3], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(whole.rnd_cons, want[4], rtol=1e-12)


@pytest.mark.skipif(R.load() is None, reason="baseline/_ref (the installed reference) is not present")
def test_installed_reference_agrees_with_port():
    x, w, genes, cells, proj, real, rnd = _job_inputs()
    ref = J.CpuJob(x, w, genes, cells, proj, real, rnd, kind="reference", chunk=9)
    assert ref.kind == "reference"
    ref.run()
    port = J.CpuJob(x, w, genes, cells, proj, real, rnd, kind="port", row_block=100)
    port.run()
    assert ref.row_names == port.row_names and ref.rnd_names == port.rnd_names
    assert np.array_equal(ref.scores_matrix("real"), port.scores_matrix("real"))       # bit for bit
    assert np.array_equal(ref.scores_matrix("rnd"), port.scores_matrix("rnd"))
    np.testing.assert_allclose(ref.cons, port.cons, rtol=1e-12)
    np.testing.assert_allclose(ref.rnd_cons, port.rnd_cons, rtol=1e-12)
    np.testing.assert_allclose(ref.logq, port.logq, rtol=1e-9, atol=1e-12)
    # dealt to steps (one sigs_vs_projections call per projection): same consistency values;
    # q-values need the whole matrix, so that variant does not report them
    split = J.CpuJob(x, w, genes, cells, proj, real, rnd, kind="reference", chunk=9, split_projections=True)
    for lo, hi in split.step_bounds(4):
        split.run(lo, hi)
    assert np.array_equal(split.cons, ref.cons) and np.isnan(split.logq).all()


def test_blas_threads_reports_a_count():
    with J.blas_threads(2) as n:
        assert n is None or n >= 1
