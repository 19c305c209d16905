"""Drive the UNMODIFIED reference (``/root/reference``, or its copy ``oracle/_ref`` made by ``oracle/build_ref.py``).

Used by ``oracle/make_golden.py`` to produce the committed fixtures under ``tests/golden/`` (build container) and by the CPU arm
of ``bench.py`` (``--impl reference`` / ``cpu_baseline``), the only run-time user.  The ``-m gpu`` tests and ``smoke()`` never
import it.

Recipe: matplotlib is absent from the image and ``import PyREMOT`` pulls it
in, so a no-op stub package (``oracle/_mplstub``) goes first on ``sys.path``; tolerances are injected
by rebinding the module-level name ``PyREMOT.docs.pbHomoReactor.solve_ivp``; grid sizes live in the
mutable dict ``PyREMOT.solvers.solSetting.solverSetting``; the RHS callbacks are plain functions on
the class, looked up at call time, so a spy can capture ``paramsSet``.
"""
from __future__ import annotations

import contextlib
import functools
import io
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; elsewhere (the GPU box) the byte-for-byte copy made by oracle/build_ref.py
REFERENCE_ROOT = os.environ.get("PBR_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/PyREMOT") else os.path.join(_HERE, "_ref"))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PyREMOT"))


def load():
    """Import the reference package; returns (PyREMOT, pbHomoReactor module, solverSetting dict)."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    for p in (REFERENCE_ROOT, os.path.join(_HERE, "_mplstub")):      # oracle/_ref carries its own stub; the extra one is harmless
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import PyREMOT
    import PyREMOT.docs.pbHomoReactor as H
    import PyREMOT.solvers.solSetting as ss
    return PyREMOT, H, ss.solverSetting


@contextlib.contextmanager
def quiet():
    """The reference prints a '\\r' progress bar from inside the RHS."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def run(modelInput, rtol=None, atol=None, zNo=None, tNo=None, capture=False):
    """rmtExe(modelInput) through the reference, optionally with injected tolerances / grid.

    Returns (result dict, captured) where captured = (t, y, paramsSet) of the first RHS call.
    """
    import scipy.integrate as si
    PyREMOT, H, setting = load()
    model = modelInput["model"]
    saved = (H.solve_ivp, dict(setting["N1"]), dict(setting["N2"]))
    name = "modelEquation" + model
    orig = getattr(H.PackedBedHomoReactorClass, name)
    cap = {}

    def spy(t, y, p):
        if not cap:
            cap["first"] = (t, list(y), p)
        return orig(t, y, p)

    try:
        if rtol is not None:
            H.solve_ivp = functools.partial(si.solve_ivp, rtol=rtol, atol=atol)
        if zNo is not None:
            setting[model]["zNo"] = zNo
        if tNo is not None:
            setting["N2"]["tNo"] = tNo
        if capture:
            setattr(H.PackedBedHomoReactorClass, name, spy)
        with quiet():
            res = PyREMOT.rmtExe(modelInput)
    finally:
        H.solve_ivp = saved[0]
        setting["N1"].update(saved[1])
        setting["N2"].update(saved[2])
        setattr(H.PackedBedHomoReactorClass, name, orig)
    return res, cap.get("first")


def rhs(model, y, paramsSet, t=0.5):
    """One call of the reference RHS callback (t=0.5 avoids the progress-bar branch at t in {0,.1,..})."""
    _, H, _ = load()
    f = getattr(H.PackedBedHomoReactorClass, "modelEquation" + model)
    with quiet():
        return f(t, y, paramsSet)
