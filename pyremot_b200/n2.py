"""Model "N2" (dynamic pseudo-homogeneous packed bed) behind rmtExe — host side of `runN2`
(`PyREMOT/docs/pbHomoReactor.py:3319-3704`): builds the per-instance table, runs K0 -> K4 -> K5b through the
C-ABI and assembles the reference's ``resPack = {"computation-time", "dataPack": [tNo dicts]}``.
"""
from __future__ import annotations

import time

import numpy as np


def _pack(modelInput, model, dataYs_k, y_k, t_end, zNo):
    """One interval's dict, `pbHomoReactor.py:3628-3678`. dataYs_k: [n+1, zNo]; y_k: [nv2, zNo] scaled."""
    compList = list(modelInput["feed"]["components"]["shell"])
    n = len(compList)
    iso = model.iso
    Cmax = float(np.max(np.asarray(modelInput["feed"]["concentration"], dtype=np.float64)))
    # quirk kept: the reference slices [:-1], which in the iso-thermal case drops the last SPECIES (:3638)
    cons1 = y_k[:-1]
    temp1 = y_k[-1] if not iso else np.repeat(0, zNo).reshape(zNo)
    return {
        "modelId": modelInput["model"],
        "processType": modelInput["operating-conditions"]["process-type"],
        "successStatus": True,
        "dataShape": (),                                   # np.array(dataTime[-1]).shape, :3618
        "labelList": compList + ["Temperature"],
        "indexList": [n, n + 1, n],                        # [compNo, indexPressure, indexTemp], :3364-3368
        "dataTime": t_end,
        "dataXs": np.linspace(0, 1, zNo),
        "dataYCons1": cons1,
        "dataYCons2": y_k[:n]*Cmax,
        "dataYTemp1": temp1,
        "dataYTemp2": dataYs_k[n:n + 1],
        "dataYs": dataYs_k,
    }


def run_n2(modelInput, device=0):
    from .api import _tolerances, get_model, solverSetting
    start = time.perf_counter()
    rtol, atol = _tolerances(modelInput)
    model = get_model(modelInput, device=device)
    zNo = solverSetting["N2"]["zNo"]
    tNo = solverSetting["N2"]["tNo"]
    opT = modelInput["operating-conditions"]["period"]
    for k in ("CaDe", "CaSpHeCa"):                          # read (and unused) by the reference, :3768-3770
        modelInput["reactor"][k]
    r = model.solve_n2_host(model.instance_table(modelInput), zNo, tNo, opT, rtol=rtol, atol=atol)
    if int(r["status"][0]) != 0:
        raise RuntimeError(f"N2 integration failed (status {int(r['status'][0])})")
    span = np.linspace(0, opT, tNo + 1)
    dataPack = [_pack(modelInput, model, r["dataYs"][0, k], r["snaps_y"][0, k], span[k + 1], zNo) for k in range(tNo)]
    return {"computation-time": round(time.perf_counter() - start, 3), "dataPack": dataPack}
