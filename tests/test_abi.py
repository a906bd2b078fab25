"""The C-ABI shared library loads (no GPU needed) and exports every symbol that
include/fastproject_b200.h declares; the ctypes table covers all of them."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fastproject_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from fastproject_b200 import _native
    return _native.load()


def test_header_declares_something():
    syms = declared_symbols()
    assert "fp_score_signatures" in syms and "fp_consistency" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export " + name


def test_ctypes_table_matches_header(lib):
    from fastproject_b200 import _native
    assert sorted(_native.SIGNATURES.keys()) == declared_symbols()


def test_version_and_error_text(lib):
    assert lib.fp_abi_version() == 3
    assert b"sm_100a" in lib.fp_version()
    assert isinstance(lib.fp_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # bad arguments are rejected before any CUDA call is made
    rc = lib.fp_score_signatures(99, None, None, None, 1, 1, 1, None, None, None, 0, 0, None, None, 1, None, 0, None)
    assert rc == -1 and b"unknown method" in lib.fp_last_error()
    rc = lib.fp_row_medians(None, 1, 1, 1, None, None)
    assert rc == -1


def test_missing_library_fails_loudly(monkeypatch):
    from fastproject_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libfastproject_b200.so")
    with pytest.raises(_native.NativeError):
        _native.load()


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fastproject_b200 import Signatures, synth, _native
    x, w, z, genes, cells = synth.make_expression(30, 10, seed=1)
    data = synth.ExpressionMatrix(x, w, genes, cells)
    sig = Signatures.Signature(dict((g, 1) for g in genes[:6]), False, "t", "s")
    with pytest.raises(_native.NativeError):
        Signatures.calculate_sig_scores(data, [sig], "weighted_avg", z, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastproject_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
