"""Recipe that makes the UNMODIFIED reference runnable next to the repo: ``oracle/_ref/`` (git-ignored, travels to the GPU box).

    python -m oracle.build_ref

The reference is pure Python, so "building" it is placing its package where the CPU arm of ``bench.py`` can import it on a
machine that has no ``/root/reference`` (the GPU box): ``/root/reference/PyREMOT`` is copied byte for byte to
``oracle/_ref/PyREMOT`` together with a no-op ``matplotlib`` stub (the image has no matplotlib and ``import PyREMOT`` pulls it
in).  Nothing under ``oracle/_ref`` is tracked by git, nothing of it is imported by the product package
(``pyremot_b200``); it is the measured CPU baseline (``bench.py --impl reference``, ``cpu_baseline.kind == "reference"``) and a
checker for the oracle.  Called by ``__graft_entry__.build()`` when ``/root/reference`` is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/PyREMOT"
DST = os.path.join(HERE, "_ref")


def build(verbose=True):
    """Returns True when oracle/_ref holds an importable copy of the reference package."""
    pkg = os.path.join(DST, "PyREMOT")
    if os.path.isdir(SRC):
        if os.path.isdir(pkg):
            shutil.rmtree(pkg)
        shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        stub = os.path.join(DST, "matplotlib")
        os.makedirs(stub, exist_ok=True)
        with open(os.path.join(stub, "__init__.py"), "w") as f:
            f.write("# no-op stand-in: the image has no matplotlib and the reference imports it at package import\n")
        shutil.copyfile(os.path.join(HERE, "_mplstub", "matplotlib", "pyplot.py"), os.path.join(stub, "pyplot.py"))
        if verbose:
            n = sum(len(fs) for _, _, fs in os.walk(pkg))
            print(f"oracle/_ref: copied {n} files of the reference package from {SRC}")
    elif verbose:
        print(f"oracle/_ref: {SRC} not present here; {'using the existing copy' if os.path.isdir(pkg) else 'NO reference copy available'}")
    return os.path.isdir(pkg)


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
