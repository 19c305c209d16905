"""Drop-in boundary: ``rmtExe(modelInput)`` / ``rmtCom()`` with PyREMOT's signatures, plus the ensemble entry.

Mirrors `PyREMOT/rmt.py:21-92` for the models this package accelerates ("N1", "N2"): same ``modelInput``
schema, same result layout (``dataPack`` / ``resPack``, `pbHomoReactor.py:2991-3007`,
`:3664-3696`), same error behaviour (prints then re-raises; unknown component ->
``Exception("Component database is not up to date!")``; missing key -> ``KeyError``; solver failure ->
``RuntimeError``).  Everything below this call runs on the GPU through libpbr.so; there is no CPU path.

Differences that are deliberate and documented in DESIGN.md:
* the integrator is the batched BDF of libpbr at tight tolerances (default rtol 1e-9 / atol 1e-13), so results
  are the *converged* solution of the reference's equations; the reference itself integrates at SciPy's
  default rtol 1e-3 and differs from its own converged answer by up to 1e-3.  ``solver-config`` accepts the
  optional keys ``rtol`` / ``atol``; ``ivp`` is accepted and ignored ("AM" is rejected: explicit fixed step).
* ``display-result`` must be "False" (no plotting in a batch solver).
* grid sizes come from ``pyremot_b200.solverSetting`` (same keys as `PyREMOT/solvers/solSetting.py:30-39`).
"""
from __future__ import annotations

import time

import numpy as np

from . import components as DB
from . import engine

# same mutable-module-dict mechanism as `PyREMOT/solvers/solSetting.py:30-39`
solverSetting = {
    "N1": {"zNo": 100},
    "N2": {"zNo": 20, "rNo": 5, "tNo": 5, "timesNo": 5},
    "M9": {"zNo": 30},            # number of output points of model M7 (`PyREMOT/docs/pbReactor.py:1283`)
    "S3": {"timesNo": 25},        # number of output points of model M8 (`PyREMOT/docs/pbReactor.py:1690`)
    "S2": {"zNo": 100, "tNo": 10, "timesNo": 5},   # grid / intervals of the dynamic model M9 (`PyREMOT/docs/pbReactor.py:2056`, `:2141-2145`)
}

DEFAULT_RTOL = 1e-9
DEFAULT_ATOL = 1e-13

_models = {}          # compiled models, most recently used last; bounded (each holds a CUDA module and scratch)
_MAX_MODELS = 16


def rmtCom():
    """Components available in the database, `PyREMOT/rmt.py:83-92`."""
    return ",".join(DB.componentSymbolList)


def _component_list(FeCom):
    """`PyREMOT/docs/rmtUtility.py:313-340`: shell + tube + medium."""
    out = []
    for k in ("shell", "tube", "medium"):
        v = FeCom.get(k)
        if v:
            out.extend(v)
    return out


def get_model(modelInput, params=(), device=0):
    """Compile (or fetch) the model for the static part of ``modelInput``."""
    probe = engine.ratec.compile_rates(modelInput["reaction-rates"], len(modelInput["feed"]["components"]["shell"]),
                                       params=params)
    key = (tuple(modelInput["feed"]["components"]["shell"]), tuple(modelInput["reactions"].items()),
           modelInput["operating-conditions"]["process-type"], probe.key, tuple(params), device,
           modelInput.get("model") if modelInput.get("model") in ("M7", "M8", "M9") else "N")
    m = _models.pop(key, None)
    if m is None:
        m = engine.Model(modelInput, params=params, device=device)
        while len(_models) >= _MAX_MODELS:                  # least recently used first
            _models.pop(next(iter(_models))).close()
    _models[key] = m
    return m


def _tolerances(modelInput):
    sc = modelInput["solver-config"]
    ivp = sc["ivp"]
    if ivp == "AM":
        raise NotImplementedError("ivp='AM' (explicit fixed-step, PyREMOT/solvers/odeSolver.py) is out of scope")
    if sc["display-result"] == "True":
        raise NotImplementedError("display-result must be \"False\": plotting is out of scope for the batched solver")
    return float(sc.get("rtol", DEFAULT_RTOL)), float(sc.get("atol", DEFAULT_ATOL))


def _n1_pack(modelInput, model, dataYs, y_scaled, elapsed):
    """One reference-layout result dict, `pbHomoReactor.py:2949-3007`.
    dataYs: [nv1, K] un-scaled; y_scaled: [nv1, K]."""
    compList = list(modelInput["feed"]["components"]["shell"])
    n = len(compList)
    iso = model.iso
    K = dataYs.shape[1]
    labelList = compList + ["Pressure"] + ([] if iso else ["Temperature"])
    Cmax = float(np.max(np.asarray(modelInput["feed"]["concentration"], dtype=np.float64)))
    cons1 = y_scaled[:n]
    theta = y_scaled[n + 1] if not iso else np.repeat(0, K).reshape(K)
    T = dataYs[n + 1:n + 2] if not iso else np.full((1, K), float(modelInput["operating-conditions"]["temperature"]))
    return {
        "modelId": modelInput["model"],
        "processType": modelInput["operating-conditions"]["process-type"],
        "successStatus": True,
        "computation-time": elapsed,
        "dataShape": (K,),
        "labelList": labelList,
        "indexList": [n, n, n + 1],
        "dataTime": [],
        "dataXs": np.linspace(0, 1, K),
        "dataYCons1": cons1,
        "dataYCons2": cons1*Cmax,
        "dataYTemp1": theta,
        "dataYTemp2": T,
        "dataYs": dataYs,
    }


def _run_n1(modelInput, device=0):
    import torch
    start = time.perf_counter()
    rtol, atol = _tolerances(modelInput)
    model = get_model(modelInput, device=device)
    K = solverSetting["N1"]["zNo"] + 1                       # `pbHomoReactor.py:2858-2862`
    tab = torch.from_numpy(model.instance_table(modelInput)).to(model._dev())
    par = model.prolog(tab)
    res = model.solve_n1(par, rtol=rtol, atol=atol, n_out=K, profile=True)
    st = int(res["status"][0].item())
    if st != 0:
        print('ODE Error')
        raise RuntimeError(f"N1 integration failed (status {st})")     # reference: bare raise -> RuntimeError
    dataYs = model.unscale_n1(par, res["prof_y"])[0].cpu().numpy()
    y_scaled = res["prof_y"][:, :, 0].t().contiguous().cpu().numpy()
    elapsed = round(time.perf_counter() - start, 3)
    return [_n1_pack(modelInput, model, dataYs, y_scaled, elapsed)]


def _run_m7(modelInput, device=0):
    """Model "M7" (`PyREMOT/docs/rmtCore.py:300-309` -> `pbReactor.py` runM3 `:1170-1369`): dimensional steady-state balances for
    C_i, T, P over z in [0, ReLe]; result ``{"dataYs": [x_i ; T] (n+1) x K, "XYList", "dataList"}`` with K =
    ``solverSetting['M9']['zNo']`` (`:1283`) points (`:1305-1367`; the pressure row is not reported).  The reference also plots
    unconditionally (`:1344`); this does not."""
    import torch
    rtol, atol = _tolerances(modelInput)
    model = get_model(modelInput, device=device)
    K = solverSetting["M9"]["zNo"]
    tab = torch.from_numpy(model.instance_table(modelInput)).to(model._dev())
    par = model.prolog(tab)
    res = model.solve_n1(par, rtol=rtol, atol=atol, n_out=K, profile=True)
    st = int(res["status"][0].item())
    if st != 0:
        raise RuntimeError(f"M7 integration failed (status {st})")
    full = model.unscale_n1(par, res["prof_y"])[0].cpu().numpy()           # [x_i ; P ; T] x K
    compList = list(modelInput["feed"]["components"]["shell"])
    n = len(compList)
    dataYs = np.concatenate([full[:n], full[n + 1:n + 2]], axis=0)
    X = np.linspace(0, modelInput["reactor"]["ReLe"], K)
    labelList = compList + ["Temperature", "Pressure"]
    XYList = [[X, row] for row in dataYs]                                   # `PyREMOT/library/plot.py:85-90`
    dataList = [{"x": X, "y": row, "leg": labelList[i]} for i, row in enumerate(dataYs)]      # `:93-115`
    return {"dataYs": dataYs, "XYList": XYList, "dataList": dataList}


def _m8_input(modelInput):
    """Model "M8" integrates the reference's BUILT-IN methanol/DME kinetics (`pbReactor.py:1903`, `modelReactions` `:8850-8930`,
    times 1000: kmol -> mol), not ``modelInput['reaction-rates']``: the compiled model gets exactly those."""
    from . import cases
    comps = list(modelInput["feed"]["components"]["shell"])
    if comps != cases.MEOH_DME_COMPONENTS:
        raise Exception("model M8 uses the reference's built-in methanol/DME kinetics: components must be " +
                        ",".join(cases.MEOH_DME_COMPONENTS))
    mi = dict(modelInput)
    v, r = cases.methanol_dme_kinetics(modelInput["reactor"]["CaBeDe"])
    mi["reaction-rates"] = {"VARS": v, "RATES": r}
    return mi


def _run_m8(modelInput, device=0):
    """Model "M8" (`PyREMOT/docs/rmtCore.py:311-320` -> `pbReactor.py` runM4 `:1580-1776`): C_i, T, P and the superficial velocity
    over z in [0, ReLe]; result ``{"dataYs": [x_i ; T] (n+1) x K, "XYList" / "dataList": the n+3 series [x_i ; T ; P ; v]}`` with
    K = ``solverSetting['S3']['timesNo']`` (`:1690`) points (`:1706-1760`)."""
    import torch
    rtol, atol = _tolerances(modelInput)
    mi = _m8_input(modelInput)
    model = get_model(mi, device=device)
    K = solverSetting["S3"]["timesNo"]
    tab = torch.from_numpy(model.instance_table(mi)).to(model._dev())
    par = model.prolog(tab)
    res = model.solve_n1(par, rtol=rtol, atol=atol, n_out=K, profile=True)
    st = int(res["status"][0].item())
    if st != 0:
        raise RuntimeError(f"M8 integration failed (status {st})")
    full = model.unscale_n1(par, res["prof_y"])[0].cpu().numpy()           # [x_i ; P ; T ; v] x K
    compList = list(modelInput["feed"]["components"]["shell"])
    n = len(compList)
    series = np.concatenate([full[:n], full[n + 1:n + 2], full[n:n + 1], full[n + 2:n + 3]], axis=0)    # [x_i ; T ; P ; v]
    X = np.linspace(0, modelInput["reactor"]["ReLe"], K)
    labelList = compList + ["Temperature", "Pressure", "Velocity"]
    XYList = [[X, row] for row in series]
    dataList = [{"x": X, "y": row, "leg": labelList[i]} for i, row in enumerate(series)]
    return {"dataYs": series[:n + 1], "XYList": XYList, "dataList": dataList}


def _run_m9(modelInput, device=0):
    """Model "M9" (`PyREMOT/docs/rmtCore.py:322-331` -> `pbReactor.py` runM5 `:1997-2294`): the dynamic dimensional model, C_i and T
    on ``solverSetting['S2']['zNo']`` nodes over ``tNo`` intervals of ``period``, pressure and velocity marched along the bed inside
    every evaluation.  The reference returns what its last plotting loop leaves behind (`:2258-2292`): the TEMPERATURE profile at the
    end of every interval, ``XYList`` = [[z, T(z, t_k)]], ``dataList`` = the same with the legend "Temperature at t=<t_k>".
    ``dataYs`` [tNo, n+1, zNo] = [x_i ; T] of every interval end is added (the reference computes it and drops it)."""
    rtol, atol = _tolerances(modelInput)
    model = get_model(modelInput, device=device)
    s2 = solverSetting["S2"]
    zNo, tNo = s2["zNo"], s2["tNo"]
    opT = modelInput["operating-conditions"]["period"]
    r = model.solve_n2_host(model.instance_table(modelInput), zNo, tNo, opT, rtol=rtol, atol=atol)
    st = int(r["status"][0])
    if st != 0:
        raise RuntimeError(f"M9 integration failed (status {st})")
    data = r["dataYs"][0]                                                     # [tNo, n+1, zNo]
    X = np.linspace(0, modelInput["reactor"]["ReLe"], zNo)
    opTSpan = np.linspace(0, opT, tNo + 1)
    XYList = [[X, data[k, -1]] for k in range(tNo)]
    dataList = [{"x": X, "y": data[k, -1], "leg": "Temperature at t=" + str(opTSpan[k + 1])} for k in range(tNo)]
    return {"XYList": XYList, "dataList": dataList, "dataYs": data}


def rmtExe(modelInput, device=0):
    """Check the model input, then run it — same contract as `PyREMOT/rmt.py:21-80`."""
    try:
        tic = time.perf_counter()
        modelType = modelInput['model']
        compList = _component_list(modelInput['feed']['components'])
        for c in compList:
            if c not in DB.componentSymbolList:
                raise Exception("Component database is not up to date!")
        if modelType == "N1":
            resModel = _run_n1(modelInput, device=device)
        elif modelType == "N2":
            from .n2 import run_n2
            resModel = run_n2(modelInput, device=device)
        elif modelType == "M7":
            resModel = _run_m7(modelInput, device=device)
        elif modelType == "M8":
            resModel = _run_m8(modelInput, device=device)
        elif modelType == "M9":
            resModel = _run_m9(modelInput, device=device)
        else:
            raise NotImplementedError(f"model {modelType!r}: only the pseudo-homogeneous packed-bed models "
                                      "N1, N2 and the dimensional models M7, M8 (steady) and M9 (dynamic) are implemented by pyremot_b200")
        # "resModel" / "comTime" are the reference's two keys (`rmt.py:73-77`).  "solver" is added so that the result itself says
        # what integrated it: the reference runs SciPy at rtol 1e-3 (its `ivp` choice), this returns the converged solution.
        sc = modelInput["solver-config"]
        return {"resModel": resModel, "comTime": (time.perf_counter() - tic)*1000,
                "solver": {"integrator": "libpbr batched BDF(1-5)/Newton, analytic Jacobian (sm_100a)",
                           "rtol": float(sc.get("rtol", DEFAULT_RTOL)), "atol": float(sc.get("atol", DEFAULT_ATOL)),
                           "ivp_requested": sc["ivp"], "ivp_honoured": False}}
    except Exception as e:
        print(e)
        raise


def rmtExeBatch(modelInput, sweep, params=(), device=0, rtol=None, atol=None, profile=False, n_out=None,
                return_device=False, strict=False):
    """Ensemble entry (new): solve N reactor instances that share ``modelInput``'s static part.

    ``sweep``: dict of per-instance arrays (see ``engine.Model.instance_table``).
    N1: returns dict(outlet [N, nv1] = last column of each instance's dataYs, dataYs [N, nv1, K] if profile,
    status [N], stats [4, N]).  N2: dict(dataYs [N, tNo, n+1, zNo], snaps_y (scaled), status, stats); grid sizes from
    ``solverSetting['N2']``.  numpy arrays (torch CUDA tensors if ``return_device``, N1 only).
    Per-instance failures are reported in ``status`` (engine codes PBR_INST_*); ``strict=True`` raises instead.
    """
    import torch
    r0, a0 = _tolerances(modelInput)
    rtol = r0 if rtol is None else rtol
    atol = a0 if atol is None else atol
    if modelInput["model"] == "M8":
        modelInput = _m8_input(modelInput)              # the reference's built-in kinetics, not modelInput['reaction-rates']
    model = get_model(modelInput, params=params, device=device)
    tab = model.instance_table(modelInput, sweep)
    if modelInput["model"] in ("N2", "M9"):
        s2 = solverSetting["N2" if modelInput["model"] == "N2" else "S2"]
        r = model.solve_n2_host(tab, s2["zNo"], s2["tNo"], modelInput["operating-conditions"]["period"], rtol=rtol, atol=atol)
    elif modelInput["model"] in ("N1", "M7", "M8"):
        K = n_out or {"N1": solverSetting["N1"]["zNo"] + 1, "M7": solverSetting["M9"]["zNo"], "M8": solverSetting["S3"]["timesNo"]}[modelInput["model"]]
        if not return_device:
            r = model.solve_n1_host(tab, rtol=rtol, atol=atol, n_out=K, profile=profile)
        else:
            tab = torch.from_numpy(tab).to(model._dev())
            par = model.prolog(tab)
            res = model.solve_n1(par, rtol=rtol, atol=atol, n_out=K, profile=profile)
            r = {"outlet": model.unscale_n1(par, res["outlet_y"][None])[:, :, 0], "status": res["status"], "stats": res["stats"]}
            if profile:
                r["dataYs"] = model.unscale_n1(par, res["prof_y"])
    else:
        raise NotImplementedError("rmtExeBatch: models N1 and N2 only")
    if strict:
        bad = int((r["status"] != 0).sum())
        if bad:
            raise RuntimeError(f"{bad} of {len(tab)} reactor instances failed to integrate")
    return r
