# -*- coding: utf-8 -*-
"""fastproject_b200 -- B200-native scoring core of FastProject.

Drop-in for the reference's data-parallel hot path only:

    from fastproject_b200 import SigScoreMethods, Signatures

mirrors ``FastProject.SigScoreMethods`` / ``FastProject.Signatures`` (scoring
half).  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C
ABI of ``include/fastproject_b200.h``; there is no CPU fallback.
"""
from . import Global  # noqa: F401  (seeds the global ``random`` stream like the reference)

__version__ = "0.1.0"
__all__ = ["SigScoreMethods", "Signatures", "NormalizationMethods", "Transforms", "Interactive", "FileIO",
           "Pipelines", "Global", "synth"]


def prefetch(*arrays, **kw):
    """Declare host expression / weight / zero-mask matrices resident and start their upload on
    a side stream; ``Signatures.calculate_sig_scores`` on the same array objects then uses the
    device copies (until ``evict``).  Overlaps the upload of the next data set with the kernels
    of the current one.  Nothing is kept resident without such a declaration."""
    from . import _engine
    _engine.prefetch(*arrays, **kw)


def evict(*arrays):
    """Release the resident device copies declared with ``prefetch``."""
    from . import _engine
    _engine.evict(*arrays)


def resident(*arrays, **kw):
    """Context manager: the matrices stay in HBM for every scoring call inside the block."""
    from . import _engine
    return _engine.resident(*arrays, **kw)


def __getattr__(name):
    # SigScoreMethods / Signatures import torch; keep ``import fastproject_b200`` light
    if name in ("SigScoreMethods", "Signatures", "NormalizationMethods", "Transforms", "Interactive", "FileIO",
                "Pipelines", "synth", "distributed"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
