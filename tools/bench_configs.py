"""Timings of the remaining BASELINE configs on one GPU (configs[1] is bench.py itself):
  config 3: N2, zNo=100, 4096 reactors        config 4 (per-GPU slice, reduced count): N2, zNo=400
  config 5: N1 finite-difference sensitivities, 16,384 base reactors x 7 variants = 114,688 solves
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases, estimation  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402


def n2(N, zNo):
    mi = cases.readme_case("N2")
    m = get_model(mi)
    rows = cases.ensemble_rows(N)
    tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
    par = m.prolog(tab)
    m.solve_n2(par, zNo, 1, 0.02)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m.solve_n2(par, zNo, 5, 0.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"N": N, "zNo": zNo, "tNo": 5, "period": 0.5, "ms": ms, "solves_per_s": N/(ms*1e-3), "failed": int((r["status"] != 0).sum()),
            "mean_steps": float(r["stats"][0].float().mean())}


def config5():
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    rows = cases.ensemble_rows(16384)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    estimation.outlet_jacobian_fd(mi, {k: v[:256] for k, v in sweep.items()}, theta)      # compile + warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    J, base = estimation.outlet_jacobian_fd(mi, sweep, theta)
    dt = time.perf_counter() - t0
    return {"base_reactors": 16384, "solves": 16384*7, "rtol": 1e-11, "seconds_end_to_end_host_buffers": dt,
            "solves_per_s": 16384*7/dt, "J_shape": list(J.shape), "finite": bool(np.isfinite(J).all())}


out = {"config3": n2(4096, 100), "config4_slice": n2(1480, 400), "config5": config5()}
print(json.dumps(out))
