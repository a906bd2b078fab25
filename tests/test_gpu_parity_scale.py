"""Parity with the oracle AT THE BENCHED SHAPES (through the Python mirror and the C ABI).

test_gpu_parity.py compares with the oracle up to 3 001 cells.  Here:

* 20 000 cells (BASELINE configs[1]: two value digits, one weight slab, 157 cell tiles): the
  whole sigs_vs_projections result against the row-blocked oracle -- consistency within 1e-6
  relative (measured ~1e-9), identical ranking of the signature x projection pairs by log10 q,
  identical significance calls; scores and ranks bit for bit;
* 100 000 cells (configs[2]: three value digits, ten weight slabs, seven j chunks): the
  contraction on 2 048 cells spread over every slab against the row-blocked oracle, the FP64
  verifier kernel on the same cells, and the medians of the tensor-core path against the FP64
  kernel's (the host cannot form all 10^10 kernel entries in test time);
* the multi-GPU job: two ranks (gloo on ONE GPU, or NCCL when two GPUs are visible) return, bit
  for bit, what one rank returns.
"""
import os
import random
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import fp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fp():
    import __graft_entry__ as ge
    ge.build()
    from fastproject_b200 import Signatures, SigScoreMethods, synth, _engine
    import types
    return types.SimpleNamespace(S=Signatures, M=SigScoreMethods, synth=synth, E=_engine)


def _conv(sigs):
    return [O.Signature(s.sig_dict, s.signed, s.source, s.name) for s in sigs]


@pytest.mark.timeout(900)
def test_config2_cell_count_whole_result_vs_row_blocked_oracle(fp):
    n, g = 20000, 1500
    proj, grad = fp.synth.make_projections(n, 2, seed=41)
    x, w, z, genes, cells = fp.synth.make_expression(g, n, seed=42, planted=grad)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    defs = fp.synth.make_signature_dicts(genes, 72, seed=43)
    sigs = [fp.S.Signature(d, sg, "t", nm) for nm, d, sg in defs]
    random.seed(9); np.random.seed(9)
    rnd = fp.S.generate_random_sigs(genes, False, [5, 10, 20, 50, 100, 200], 40)     # 240 -> 312 rows: 2 CTA-pair tiles
    with fp.E.resident(x, w):
        real = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", None, 5)
        bg = fp.S.calculate_sig_scores(data, rnd, "weighted_avg", None, 5)
    got = fp.S.sigs_vs_projections(proj, real, bg, precision="tc")

    odata = O.ExpressionMatrix(x, w, genes, cells)
    oreal = O.calculate_sig_scores(odata, _conv(sigs), "weighted_avg", None, 5)
    obg = O.calculate_sig_scores(odata, _conv(rnd), "weighted_avg", None, 5)
    for k in oreal:
        assert np.array_equal(real[k].scores, oreal[k].scores), k
    some = list(obg.keys())[::17]
    for k in some:
        assert np.array_equal(bg[k].scores, obg[k].scores), k
        assert np.array_equal(bg[k].ranks, obg[k].ranks), k
    want = O.sigs_vs_projections(proj, oreal, obg, row_block=500)
    assert got[0] == want[0] and got[1] == want[1] and got[5] == want[5]
    np.testing.assert_allclose(got[2], want[2], rtol=1e-6)           # the north-star tolerance ...
    np.testing.assert_allclose(got[2], want[2], rtol=2e-8)           # ... and what the digit planes give
    np.testing.assert_allclose(got[4], want[4], rtol=2e-8)
    np.testing.assert_allclose(got[3], want[3], rtol=1e-6, atol=1e-6)
    assert np.array_equal(np.argsort(got[3], axis=None, kind="stable"),
                          np.argsort(want[3], axis=None, kind="stable"))
    assert np.array_equal(got[3] < -1.3, want[3] < -1.3)
    assert (want[3] < -1.3).any() and not (want[3] < -1.3).all()


@pytest.mark.timeout(900)
def test_config3_cell_count_three_digit_path_vs_row_blocked_oracle(fp):
    import torch
    from scipy.stats import rankdata
    n, k = 100000, 70
    rng = np.random.RandomState(51)
    proj, _ = fp.synth.make_projections(n, 1, seed=52)
    coords = proj["Proj00"]
    vals = np.array([rankdata(rng.normal(size=n) + (2.0 * coords[0] if r % 3 == 0 else 0)) for r in range(k)])
    vals[k - 1] = rankdata(np.round(rng.normal(size=n), 2))             # ties -> half-integer ranks
    ld = fp.E.padded(n)
    dev = torch.zeros((k, ld), dtype=torch.float64, device="cuda")
    dev[:, :n] = torch.from_numpy(vals).cuda()
    c = torch.from_numpy(coords).cuda()
    tc = fp.E.consistency(c, 0.33 ** 2, dev, n, precision="tc", slot="scale")
    assert not fp.E.consistency_refused(dev, n, precision="tc", slot="scale")
    f64 = fp.E.consistency(c, 0.33 ** 2, dev[:8], n, precision="fp64")
    # cells spread over every weight slab and every part of a cell tile
    cells = np.sort(rng.choice(n, size=2048, replace=False))
    cells[:4] = [0, 1, n // 2, n - 1]
    cells = np.unique(cells)
    got = tc.index_select(1, torch.from_numpy(cells).cuda()).cpu().numpy()
    got64 = f64.index_select(1, torch.from_numpy(cells).cuda()).cpu().numpy()
    vt = np.ascontiguousarray(vals.T)
    want = np.empty((k, cells.size))
    for b in range(0, cells.size, 256):
        sel = cells[b:b + 256]
        wt = O.neighbourhood_weights(coords, 0.33, sel)
        want[:, b:b + 256] = np.abs(vt[sel] - np.dot(wt, vt)).T
    # |v - prediction| in rank units; 1e-6 relative on consistency = 1 - med/N allows 0.075 here
    assert np.abs(got - want).max() <= 1e-7 * n
    assert np.abs(got64 - want[:8]).max() <= 1e-9 * n
    # medians over ALL cells: tensor-core path against the FP64 kernel (itself checked above)
    med_tc = fp.E.row_medians(tc[:8], n).cpu().numpy()
    med_64 = fp.E.row_medians(f64, n).cpu().numpy()
    np.testing.assert_allclose(1 - med_tc / n, 1 - med_64 / n, rtol=1e-7)
    assert np.array_equal(np.argsort(med_tc, kind="stable"), np.argsort(med_64, kind="stable"))


_TWO_RANK_SCRIPT = r"""
import io, contextlib, os, random, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from fastproject_b200 import Signatures as S, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank % ngpu)
backend = "nccl" if ngpu >= world else "gloo"          # two ranks on one GPU: gloo, staged through the host
dist.init_process_group(backend)
n, g = 1500, 500
proj, grad = synth.make_projections(n, 3, seed=61)
x, w, z, genes, cells = synth.make_expression(g, n, seed=62, planted=grad)
data = synth.ExpressionMatrix(x, w, genes, cells)
defs = synth.make_signature_dicts(genes, 301, seed=63)
sigs = [S.Signature(d, sg, "t", nm) for nm, d, sg in defs]
sigs[7] = S.Signature(dict(NOPE=1), False, "t", "DROPPED")        # rejected: uneven row blocks
random.seed(4); np.random.seed(4)
rnd = S.generate_random_sigs(genes, False, [5, 10, 20, 50, 100, 200], 45)
from fastproject_b200.SigScoreMethods import SignatureScores
pre = SignatureScores(np.asarray(grad), "PRE_NUM", cells, isFactor=False, isPrecomputed=True, numGenes=0)
with contextlib.redirect_stdout(io.StringIO()):
    real = S.calculate_sig_scores(data, sigs, "weighted_avg", z, 5, distributed=True)
    bg = S.calculate_sig_scores(data, rnd, "weighted_avg", z, 5, distributed=True)
    real["PRE_NUM"] = pre
    out = S.sigs_vs_projections(proj, real, bg, precision={precision!r}, distributed=True)
scores = np.array([real[k].scores for k in list(real.keys())[:40]])       # gathers (collective)
if rank == 0:
    np.savez({out!r}, names=np.array(out[0]), cols=np.array(out[1]), cons=out[2], logq=out[3], rnd=out[4],
             rnd_names=np.array(out[5]), scores=scores, backend=np.array(backend))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.timeout(900)
@pytest.mark.parametrize("precision", ["tc", "fp64"])
def test_two_rank_job_equals_one_rank_job(fp, precision, tmp_path):
    """sigs_vs_projections(distributed=True) on two ranks == the one-rank result, bit for bit
    (the partition changes who computes a (projection, signature) pair, not how)."""
    import io
    import contextlib
    out = str(tmp_path / "two_rank.npz")
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK_SCRIPT.format(root=ROOT, precision=precision, out=out))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, PYTHONPATH=ROOT)
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=800)
    assert proc.returncode == 0, proc.stdout[-4000:]
    got = np.load(out)
    # the same job on one rank, in this process
    n, g = 1500, 500
    proj, grad = fp.synth.make_projections(n, 3, seed=61)
    x, w, z, genes, cells = fp.synth.make_expression(g, n, seed=62, planted=grad)
    data = fp.synth.ExpressionMatrix(x, w, genes, cells)
    defs = fp.synth.make_signature_dicts(genes, 301, seed=63)
    sigs = [fp.S.Signature(d, sg, "t", nm) for nm, d, sg in defs]
    sigs[7] = fp.S.Signature(dict(NOPE=1), False, "t", "DROPPED")
    random.seed(4); np.random.seed(4)
    rnd = fp.S.generate_random_sigs(genes, False, [5, 10, 20, 50, 100, 200], 45)
    pre = fp.M.SignatureScores(np.asarray(grad), "PRE_NUM", cells, isFactor=False, isPrecomputed=True, numGenes=0)
    with contextlib.redirect_stdout(io.StringIO()):
        real = fp.S.calculate_sig_scores(data, sigs, "weighted_avg", z, 5)
        bg = fp.S.calculate_sig_scores(data, rnd, "weighted_avg", z, 5)
        real["PRE_NUM"] = pre
        want = fp.S.sigs_vs_projections(proj, real, bg, precision=precision)
    assert "DROPPED" not in want[0] and want[0][-1] == "PRE_NUM"
    assert list(got["names"]) == want[0] and list(got["cols"]) == want[1] and list(got["rnd_names"]) == want[5]
    assert np.array_equal(got["cons"], want[2])
    assert np.array_equal(got["logq"], want[3])
    assert np.array_equal(got["rnd"], want[4])
    assert np.array_equal(got["scores"], np.array([real[k].scores for k in list(real.keys())[:40]]))
