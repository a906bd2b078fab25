# -*- coding: utf-8 -*-
"""The UNMODIFIED reference as a travelling checker and CPU baseline.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python.  ``install()`` runs the base contract's one offline install,

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(from a copy under /tmp because setuptools writes build/ and egg-info into the source tree and
``/root/reference`` is read-only; ``--no-deps`` because numpy / scipy / pandas / scikit-learn of
this image are used as they are).  ``baseline/_ref`` is git-ignored -- reference sources never
enter this repository's history -- but not gpurun-ignored, so the installed package travels to
the GPU box like the built ``.so``.  ``__graft_entry__.build()`` calls ``install()`` in the
authoring container; on the GPU box ``/root/reference`` does not exist and the installed copy is
used as found.

``load()`` imports that copy (``FastProject.Signatures`` / ``FastProject.SigScoreMethods`` import
unmodified on this image) and returns the two hot-path modules, or None when
it is absent -- callers then use the restatement in ``oracle/fp_oracle.py`` and say so
(``kind: "port"``).

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DST = os.path.join(ROOT, "baseline", "_ref")


def installed():
    return os.path.isfile(os.path.join(REF_DST, "FastProject", "Signatures.py"))


def install(verbose=False):
    """Install the reference into baseline/_ref (once).  Returns True if it is there after the
    call; never raises (an uninstallable reference just means ``kind: "port"``)."""
    if installed() or not os.path.isdir(REF_SRC):
        return installed()
    tmp = tempfile.mkdtemp(prefix="fp_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_DST, src]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=tmp)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout[-2000:])
    except Exception as exc:                       # pragma: no cover
        sys.stderr.write("reference install failed: %r\n" % (exc,))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


_loaded = None


def load():
    """(FastProject.Signatures, FastProject.SigScoreMethods) of the installed copy, or None."""
    global _loaded
    if _loaded is not None:
        return _loaded or None
    if not installed():
        _loaded = False
        return None
    sys.dont_write_bytecode = True
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    try:
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            from FastProject import Signatures as ref_sig
            from FastProject import SigScoreMethods as ref_ssm
    except Exception as exc:                       # pragma: no cover - depends on the image
        sys.stderr.write("baseline/_ref present but not importable: %r\n" % (exc,))
        _loaded = False
        return None
    _loaded = (ref_sig, ref_ssm)
    return _loaded
