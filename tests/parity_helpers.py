"""Oracle solves for the GPU parity tests, spread over the host cores (test infrastructure).

Each worker is a plain subprocess of this file (no multiprocessing start-method subtleties next to a CUDA context): it rebuilds
``modelInput`` from the numeric rows of its job file (the rate lambdas cannot be pickled) and integrates the NumPy restatement
of the reference RHS with SciPy at converged tolerances (`oracle/pbr_oracle.py`).
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mi(kind, theta_as_vars=False):
    from pyremot_b200 import cases
    if kind == "N1":
        return cases.readme_case("N1", theta_as_vars=theta_as_vars)
    if kind == "N1iso":
        return cases.readme_case("N1", "iso-thermal")
    return cases.readme_case("N2")


def _work(job):
    """job: npz with `kind`, sweep arrays `sw_<key>`, and the solver settings; returns the result array."""
    from oracle import pbr_oracle as O
    kind = str(job["kind"])
    sweep = {k[3:]: job[k] for k in job.files if k.startswith("sw_")}
    rtol, atol = float(job["rtol"]), float(job["atol"])
    if kind in ("N1", "N1iso"):
        ens = O.Ensemble(_mi(kind, theta_as_vars=bool(job["theta"])), sweep)
        _, data, _ = O.solve_n1(ens, method="BDF", rtol=rtol, atol=atol)
        return data if bool(job["profiles"]) else data[:, :, -1]
    if kind == "M9":
        from pyremot_b200 import cases
        ens = O.Ensemble(cases.m9_case(period=float(job["period"])), sweep)
        Y, _, _ = O.solve_m9(ens, zNo=int(job["zNo"]), tNo=int(job["tNo"]), method="BDF", rtol=rtol, atol=atol, sparse=True)
        return Y                                               # dimensional [C_i ; T] snapshots
    ens = O.Ensemble(_mi("N2"), sweep)
    ens.period = float(job["period"])
    _, data, _ = O.solve_n2(ens, zNo=int(job["zNo"]), tNo=int(job["tNo"]), method="BDF", rtol=rtol, atol=atol, sparse=True)
    return data


def _run_jobs(sweep, idx, timeout, **settings):
    idx = list(idx)
    n = max(1, min(len(idx), os.cpu_count() or 1))
    parts = [idx[k::n] for k in range(n)]
    tmp = tempfile.mkdtemp(prefix="pbr_parity_")
    procs = []
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1",
               CUDA_VISIBLE_DEVICES="")
    for k, part in enumerate(parts):
        jf, of = os.path.join(tmp, f"job{k}.npz"), os.path.join(tmp, f"out{k}.npy")
        np.savez(jf, **{"sw_" + key: np.asarray(v)[part] for key, v in sweep.items()}, **settings)
        procs.append((part, of, subprocess.Popen([sys.executable, os.path.abspath(__file__), jf, of], env=env,
                                                 stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)))
    out = {}
    for part, of, p in procs:
        try:
            _, err = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for _, _, q in procs:
                q.kill()
            raise
        if p.returncode != 0:
            raise RuntimeError("oracle worker failed:\n" + err.decode()[-2000:])
        res = np.load(of)
        for j, i in enumerate(part):
            out[i] = res[j]
    return np.stack([out[i] for i in idx])


def oracle_n1(sweep, idx, kind="N1", profiles=False, rtol=1e-11, atol=1e-14, theta_as_vars=False, timeout=1500):
    """Converged oracle results of rows ``idx`` of ``sweep``: outlets [len(idx), nv1] or profiles [len(idx), nv1, 101]."""
    return _run_jobs(sweep, idx, timeout, kind=kind, profiles=profiles, rtol=rtol, atol=atol, theta=theta_as_vars)


def oracle_n2(sweep, idx, zNo, tNo, period, rtol=1e-9, atol=1e-13, timeout=1500):
    """Converged oracle snapshots dataYs [len(idx), tNo, n+1, zNo] of rows ``idx``."""
    return _run_jobs(sweep, idx, timeout, kind="N2", zNo=zNo, tNo=tNo, period=period, rtol=rtol, atol=atol)


def oracle_m9(sweep, idx, zNo, tNo, period, rtol=1e-8, atol=1e-11, timeout=1500):
    """Converged oracle snapshots [len(idx), tNo, n+1, zNo] (dimensional C_i, T) of the dynamic dimensional model M9."""
    return _run_jobs(sweep, idx, timeout, kind="M9", zNo=zNo, tNo=tNo, period=period, rtol=rtol, atol=atol)


def relerr_rows(a, b, floor=0.0):
    """Max relative error per leading row."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    e = np.abs(a - b)/(np.abs(b) + floor)
    return e.reshape(e.shape[0], -1).max(axis=1)


if __name__ == "__main__":
    np.save(sys.argv[2], _work(np.load(sys.argv[1])))
