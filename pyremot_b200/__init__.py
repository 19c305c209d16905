"""pyremot_b200 — B200-native batched solver for PyREMOT's pseudo-homogeneous packed-bed path (N1, N2).

Public surface mirrors `PyREMOT/__init__.py:1-3`: ``rmtExe``, ``rmtCom`` (+ the new ensemble entry
``rmtExeBatch`` and the mutable ``solverSetting`` dict of `PyREMOT/solvers/solSetting.py`).
"""
from .api import rmtExe, rmtCom, rmtExeBatch, solverSetting  # noqa: F401

__version__ = "0.1.0"
