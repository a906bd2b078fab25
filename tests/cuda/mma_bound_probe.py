"""What bounds consistency_mma_kernel: the same launch timed with the release library, with a
build that issues the MMAs but moves no operands (-DFP_TC_TIMING_NO_LOADS) and with one that
moves the operands but issues no MMAs (-DFP_TC_TIMING_NO_MMA).  The variants are compiled by
`python tests/cuda/mma_bound_probe.py build` (here, nvcc) into tests/cuda/_bin/ and timed by
`python tests/cuda/mma_bound_probe.py <variant>` on the GPU box, one process per variant.
Their outputs are garbage by construction: only the times mean anything."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "tests", "cuda", "_bin")
VARIANTS = {"noloads": ["-DFP_TC_TIMING_NO_LOADS"], "nomma": ["-DFP_TC_TIMING_NO_MMA"],
            # L2 promotion of the TMA loads (the release build's is in consistency_tc.cu)
            "promo128": ["-DFP_TC_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_L2_128B"],
            "promonone": ["-DFP_TC_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_NONE"],
            "promo128_nomma": ["-DFP_TC_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_L2_128B", "-DFP_TC_TIMING_NO_MMA"],
            "plainarrive_nomma": ["-DFP_TC_TIMING_PLAIN_ARRIVE", "-DFP_TC_TIMING_NO_MMA"],
            "noskip_nomma": ["-DFP_TC_TIMING_NO_SKIP", "-DFP_TC_TIMING_NO_MMA"],
            "noskip": ["-DFP_TC_TIMING_NO_SKIP"],
            "spin_nomma": ["-DFP_TC_SPIN_WAITS", "-DFP_TC_TIMING_NO_MMA"],
            "spin": ["-DFP_TC_SPIN_WAITS"],
            "sametile_nomma": ["-DFP_TC_TIMING_SAME_TILE", "-DFP_TC_TIMING_NO_MMA"],
            "twostages_nomma": ["-DFP_TC_TIMING_TWO_STAGES", "-DFP_TC_TIMING_NO_MMA"],
            "twostages": ["-DFP_TC_TIMING_TWO_STAGES"],
            "promonone_nomma": ["-DFP_TC_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_NONE", "-DFP_TC_TIMING_NO_MMA"]}


def build(only=None):
    import __graft_entry__ as ge
    os.makedirs(BIN, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(ge.CSRC, "*.cu")))
    for name, flag in VARIANTS.items():
        if only and name not in only:
            continue
        objs = []
        for src in sources:
            obj = os.path.join(BIN, "%s_%s.o" % (name, os.path.basename(src)[:-3]))
            extra = flag if src.endswith("consistency_tc.cu") else []
            subprocess.check_call([ge._nvcc()] + ge.NVCC_FLAGS + extra + ["-c", src, "-o", obj])
            objs.append(obj)
        subprocess.check_call([ge._nvcc(), "-shared", "-o", os.path.join(BIN, "libfp_%s.so" % name)] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a"])
        for obj in objs:
            os.remove(obj)


def run(variant, n=20000, k=3500, reps=4):
    import torch
    from fastproject_b200 import _native
    if variant != "release":
        _native.LIB_PATH = os.path.join(BIN, "libfp_%s.so" % variant)
    from fastproject_b200 import _engine as E, synth
    proj, _ = synth.make_projections(n, 1, seed=5)
    c = torch.from_numpy(proj["Proj00"]).cuda()
    ld = E.padded(n)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    ranks = E.rank_rows(torch.rand((k, ld), dtype=torch.float64, device="cuda", generator=g), n)
    out = torch.empty_like(ranks)
    E.consistency(c, 0.33 ** 2, ranks, n, out=out, precision="tc", any_cell_order=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        E.consistency(c, 0.33 ** 2, ranks, n, out=out, precision="tc", any_cell_order=True)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
    print("%-8s n=%d k=%d: fp_consistency %s ms (all its kernels; the MMA kernel is the rest after ~1.2 ms of "
          "set-up, planes and packing)" % (variant, n, k, " ".join("%.3f" % m for m in ms)))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[1], *[int(a) for a in sys.argv[2:]])
