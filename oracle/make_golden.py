"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference from ``/root/reference``.

Run in the build container only (the reference cannot travel):

    python -m oracle.make_golden [--only props,n1_rhs,...] [--slow]

Every array stored here is an output of reference code (``rmtExe`` or its helper / RHS callbacks);
the inputs are rebuilt at test time from ``pyremot_b200.cases`` with the stored seeds / states.
Groups:
  props      known-answer property values (dH25, Cp-mean, dH correction, viscosities, Wilke)
  n1_rhs     modelEquationN1 on random states (README case non-iso + iso; swept inputs)
  n2_rhs     modelEquationN2 on random states (README N2, tests/test_rmt_N1_DME.py, CH4 case)
  n1_solve   rmtExe N1: default LSODA output and converged (rtol 1e-12 / atol 1e-14) profiles
  n1_sweep   converged N1 outlets for 16 corner + 16 random rows of the config-2 sweep
  n2_solve   rmtExe N2 zNo=20, tNo=5: default BDF and tightened (rtol 1e-9) snapshots   [--slow]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_driver as ref  # noqa: E402
from pyremot_b200 import cases  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    p = os.path.join(OUT, name + ".npz")
    np.savez_compressed(p, **arrs)
    print(f"wrote {p}: " + ", ".join(f"{k}{np.shape(v)}" for k, v in arrs.items()), flush=True)


def g_props():
    ref.load()
    from PyREMOT.docs.rmtThermo import (calStandardEnthalpyOfReaction, calMeanHeatCapacityAtConstantPressure,
                                        calEnthalpyChangeOfReaction, calHeatCapacityAtConstantPressure)
    from PyREMOT.docs.gasTransPor import calGasViscosity, calMixturePropertyM1
    from PyREMOT.docs.rmtUtility import rmtUtilityClass as U
    from PyREMOT import rmtCom
    from PyREMOT.data.componentData import componentDataSource
    comps = cases.MEOH_DME_COMPONENTS
    allc = rmtCom().split(",")
    Ts = np.array([298.15, 400.0, 493.0, 523.0, 553.0, 620.8, 700.0, 973.0])
    rls = U.buildReactionCoefficient(cases.MEOH_DME_REACTIONS)
    MW_all = np.array([float(c["MW"]) for c in componentDataSource])
    x0 = np.array([574.8978, 287.4489, 1.15e-02, 287.4489, 1.15e-02, 1.15e-02])
    x0 = x0/x0.sum()
    MW = np.array([MW_all[allc.index(s)] for s in comps])
    save("props",
         symbols=np.array(allc),
         Ts=Ts,
         MW_all=MW_all,
         dH25=np.array([calStandardEnthalpyOfReaction(r) for r in cases.MEOH_DME_REACTIONS.values()]),
         dH25_ch4=np.array([calStandardEnthalpyOfReaction("2CH4 <=> C2H4 + 2H2")]),
         cp_all=np.array([calHeatCapacityAtConstantPressure(allc, T) for T in Ts]),
         cpmean_all=np.array([calMeanHeatCapacityAtConstantPressure(allc, T) for T in Ts]),
         dHcorr=np.array([calEnthalpyChangeOfReaction(rls, T) for T in Ts]),
         mu_all=np.array([calGasViscosity(allc, T) for T in Ts]),
         mu_mix=np.array([calMixturePropertyM1(6, calGasViscosity(comps, T), x0, MW) for T in Ts]),
         x0=x0)


def _n1_states(rng, iv, K):
    n = iv.size - 2
    Y = np.zeros((K, iv.size))
    Y[:, :n] = iv[:n]*np.exp(rng.normal(0, 0.5, (K, n)))*10.0**(rng.uniform(-1, 2, (K, n))*(iv[:n] < 1e-3))
    Y[:, n] = rng.uniform(0.9, 1.0, K)
    Y[:, n+1] = rng.uniform(-0.02, 0.35, K)
    return Y


def g_n1_rhs():
    rng = np.random.default_rng(101)
    K = 600
    out = {}
    for tag, pt in (("noniso", "non-iso-thermal"), ("iso", "iso-thermal")):
        mi = cases.readme_case("N1", pt)
        res, (t0, y0, ps) = ref.run(mi, capture=True)
        y0 = np.array(y0)
        full = _n1_states(rng, np.concatenate([y0, [0.0]]) if pt == "iso-thermal" else y0, K)
        Y = full[:, :y0.size]
        Y[0] = y0
        F = np.array([ref.rhs("N1", list(y), ps) for y in Y])
        out[f"y_{tag}"], out[f"f_{tag}"] = Y, F
    # swept inputs: one RHS at the inlet state + one random state for 24 rows of the config-2 table
    rows = cases.ensemble_rows(24)
    Ys, Fs = [], []
    for i in range(24):
        mi = cases.readme_case("N1", T=float(rows["T"][i]), P=float(rows["P"][i]), conc=rows["conc"][i],
                               Tm=float(rows["T"][i]))
        res, (t0, y0, ps) = ref.run(mi, capture=True)
        y1 = _n1_states(rng, np.array(y0), 1)[0]
        Ys.append([y0, list(y1)])
        Fs.append([ref.rhs("N1", list(y0), ps), ref.rhs("N1", list(y1), ps)])
    out["y_sweep"], out["f_sweep"] = np.array(Ys), np.array(Fs)
    save("n1_rhs", **out)


def _n2_states(rng, iv2d, K, neg=True):
    nv, zNo = iv2d.shape
    n = nv - 1
    Y = np.zeros((K, nv, zNo))
    base = iv2d[:n, :1]
    Y[:, :n, :] = base*np.exp(rng.normal(0, 0.4, (K, n, zNo)))*10.0**(rng.uniform(-1, 2, (K, n, zNo))*(base < 1e-3))
    Y[:, n, :] = rng.uniform(-0.02, 0.3, (K, zNo))
    if neg:  # exercise the 1e-30 clamp (pbHomoReactor.py:3899, :4093)
        Y[1::3, 2, zNo//2] = -1e-7
        Y[2::3, n-1, 1] = -3e-6
    return Y


def g_n2_rhs():
    rng = np.random.default_rng(202)
    out = {}
    specs = [("readme_z5", cases.readme_case("N2"), 5, 12),
             ("readme_z12", cases.readme_case("N2"), 12, 12),
             ("readme_iso_z8", cases.readme_case("N2", "iso-thermal"), 8, 8),
             ("dme_z20", cases.dme_n2_case(), 20, 6),
             ("ch4_z10", cases.ch4_n2_case(), 10, 8)]
    for tag, mi, zNo, K in specs:
        mi["operating-conditions"]["period"] = 1e-9  # the run is only needed to capture paramsSet
        res, (t0, y0, ps) = ref.run(mi, zNo=zNo, tNo=1, capture=True)
        iso = mi["operating-conditions"]["process-type"] == "iso-thermal"
        y0 = np.array(y0).reshape(-1, zNo)
        iv = np.concatenate([y0, np.zeros((1, zNo))]) if iso else y0
        Y = _n2_states(rng, iv, K)[:, :y0.shape[0], :]
        Y[0] = y0
        F = np.array([np.array(ref.rhs("N2", list(y.reshape(-1)), ps)).reshape(y.shape) for y in Y])
        out[f"y_{tag}"], out[f"f_{tag}"] = Y, F
    save("n2_rhs", **out)


def _pack_n1(res):
    d = res["resModel"][0]
    return {k: np.array(d[k]) for k in ("dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2", "dataYs")}


def g_n1_solve():
    out = {}
    for tag, pt in (("noniso", "non-iso-thermal"), ("iso", "iso-thermal")):
        for ivp in ("default", "BDF", "Radau"):
            t = time.time()
            res, _ = ref.run(cases.readme_case("N1", pt, ivp=ivp))
            for k, v in _pack_n1(res).items():
                out[f"{tag}_{ivp}_{k}"] = v
            print(f"n1 {tag} {ivp} default tolerances {time.time()-t:.1f}s", flush=True)
        t = time.time()
        res, _ = ref.run(cases.readme_case("N1", pt, ivp="default"), rtol=1e-12, atol=1e-14)
        for k, v in _pack_n1(res).items():
            out[f"{tag}_conv_{k}"] = v
        res2, _ = ref.run(cases.readme_case("N1", pt, ivp="Radau"), rtol=1e-11, atol=1e-14)
        out[f"{tag}_conv_radau_dataYs"] = np.array(res2["resModel"][0]["dataYs"])
        d = res["resModel"][0]
        out[f"{tag}_labelList"] = np.array(d["labelList"])
        out[f"{tag}_indexList"] = np.array(d["indexList"])
        out[f"{tag}_dataShape"] = np.array(d["dataShape"])
        print(f"n1 {tag} converged {time.time()-t:.1f}s", flush=True)
    save("n1_solve", **out)


def sweep_corner_rows():
    c = np.array([[T, P, h, c2] for T in (493.0, 553.0) for P in (3e6, 7e6) for h in (1.0, 3.0) for c2 in (0.2, 0.8)])
    x0 = cases.feed_mole_fraction(c[:, 2], c[:, 3])
    return {"T": c[:, 0], "P": c[:, 1], "conc": x0*(c[:, 1]/(cases.R_CONST*c[:, 0]))[:, None]}


def g_n1_sweep():
    cr = sweep_corner_rows()
    rr = cases.ensemble_rows(16)
    T = np.concatenate([cr["T"], rr["T"]])
    P = np.concatenate([cr["P"], rr["P"]])
    conc = np.concatenate([cr["conc"], rr["conc"]])
    outs, profs = [], []
    for i in range(len(T)):
        t = time.time()
        mi = cases.readme_case("N1", T=float(T[i]), P=float(P[i]), conc=conc[i], Tm=float(T[i]), ivp="BDF")
        res, _ = ref.run(mi, rtol=1e-11, atol=1e-14)
        ys = np.array(res["resModel"][0]["dataYs"])
        outs.append(ys[:, -1])
        profs.append(ys)
        print(f"n1 sweep row {i} {time.time()-t:.1f}s Tout={ys[-1, -1]:.2f} Tmax={ys[-1].max():.2f}", flush=True)
    save("n1_sweep", T=T, P=P, conc=conc, outlet=np.array(outs), dataYs=np.array(profs))


def _pack_n2(res):
    dp = res["resModel"]["dataPack"]
    return {"dataTime": np.array([d["dataTime"] for d in dp]),
            "dataXs": np.array(dp[0]["dataXs"]),
            "dataYCons1": np.array([d["dataYCons1"] for d in dp]),
            "dataYCons2": np.array([d["dataYCons2"] for d in dp]),
            "dataYTemp1": np.array([d["dataYTemp1"] for d in dp]),
            "dataYTemp2": np.array([d["dataYTemp2"] for d in dp]),
            "dataYs": np.array([d["dataYs"] for d in dp])}


def g_n2_solve():
    out = {}
    specs = [("readme_z20", cases.readme_case("N2", ivp="BDF"), 20, 5),
             ("dme_z20", cases.dme_n2_case(ivp="BDF"), 20, 5),
             ("ch4_z10", cases.ch4_n2_case(ivp="BDF"), 10, 5),
             ("readme_iso_z10", cases.readme_case("N2", "iso-thermal", ivp="BDF"), 10, 5)]
    for tag, mi, zNo, tNo in specs:
        t = time.time()
        res, _ = ref.run(mi, zNo=zNo, tNo=tNo)
        for k, v in _pack_n2(res).items():
            out[f"{tag}_default_{k}"] = v
        print(f"n2 {tag} default {time.time()-t:.1f}s", flush=True)
        t = time.time()
        res, _ = ref.run(mi, rtol=1e-9, atol=1e-12, zNo=zNo, tNo=tNo)
        for k, v in _pack_n2(res).items():
            out[f"{tag}_conv_{k}"] = v
        print(f"n2 {tag} rtol 1e-9 {time.time()-t:.1f}s", flush=True)
        save("n2_solve", **out)


def g_m7():
    """Dimensional steady-state model M7 (rmtCore.py:300-309 -> pbReactor.py runM3 / modelEquationM3): RHS on random states,
    rmtExe at the default tolerances, and converged outlets / profiles for the README case and 8 swept rows."""
    import functools
    import scipy.integrate as si
    PyREMOT, H, setting = ref.load()
    import PyREMOT.docs.pbReactor as R
    cls = R.PackedBedReactorClass
    orig = cls.modelEquationM3
    cap = {}

    def spy(t, y, a, b, c):
        cap.setdefault("args", (a, b, c))
        return orig(t, y, a, b, c)

    def run(mi, rtol=None, atol=None):
        saved = R.solve_ivp
        try:
            if rtol is not None:
                R.solve_ivp = functools.partial(si.solve_ivp, rtol=rtol, atol=atol)
            with ref.quiet():
                return PyREMOT.rmtExe(mi)["resModel"]
        finally:
            R.solve_ivp = saved
    out = {}
    cls.modelEquationM3 = spy
    try:
        res = run(cases.m7_case())
    finally:
        cls.modelEquationM3 = orig
    out["default_dataYs"] = np.array(res["dataYs"])
    out["default_X"] = np.array(res["XYList"][0][0])
    out["default_legs"] = np.array([d["leg"] for d in res["dataList"]])
    rng = np.random.default_rng(707)
    mi = cases.m7_case()
    y0 = np.array(list(mi["feed"]["concentration"]) + [523.0, 5e6])
    Y = np.tile(y0, (200, 1))
    Y[:, :6] *= np.exp(rng.normal(0, 0.4, (200, 6)))*10.0**(rng.uniform(-1, 2, (200, 6))*(y0[:6] < 1.0))
    Y[:, 6] = rng.uniform(490, 700, 200)
    Y[:, 7] *= rng.uniform(0.9, 1.0, 200)
    Y[0] = y0
    with ref.quiet():
        out["rhs_y"], out["rhs_f"] = Y, np.array([orig(0.5, list(y), *cap["args"]) for y in Y])
    t = time.time()
    res = run(cases.m7_case(), rtol=1e-12, atol=1e-12)
    out["conv_dataYs"] = np.array(res["dataYs"])
    rows = cases.ensemble_rows(8)
    out["sweep_conv_dataYs"] = np.array([np.array(run(cases.m7_case(T=float(rows["T"][i]), P=float(rows["P"][i]), conc=rows["conc"][i],
                                                                   Tm=float(rows["T"][i])), rtol=1e-12, atol=1e-12)["dataYs"])
                                          for i in range(8)])
    print(f"m7 converged runs {time.time()-t:.1f}s", flush=True)
    save("m7", **out)


def g_m8():
    """Dimensional steady-state model M8 (rmtCore.py:311-320 -> pbReactor.py runM4 / modelEquationM4): RHS on random states,
    rmtExe at the default tolerances, converged profiles for the README case and 8 swept rows."""
    import functools
    import scipy.integrate as si
    PyREMOT, H, setting = ref.load()
    import PyREMOT.docs.pbReactor as R
    cls = R.PackedBedReactorClass
    orig = cls.modelEquationM4
    cap = {}

    def spy(t, y, a, b, c):
        cap.setdefault("args", (a, b, c))
        return orig(t, y, a, b, c)

    def run(mi, rtol=None, atol=None):
        saved = R.solve_ivp
        try:
            if rtol is not None:
                R.solve_ivp = functools.partial(si.solve_ivp, rtol=rtol, atol=atol)
            with ref.quiet():
                return PyREMOT.rmtExe(mi)["resModel"]
        finally:
            R.solve_ivp = saved

    def series(res):            # [x_i ; T ; P ; v] from the plotted series (dataYs itself holds only x_i and T)
        return np.array([xy[1] for xy in res["XYList"]])
    out = {}
    cls.modelEquationM4 = spy
    try:
        res = run(cases.m8_case())
    finally:
        cls.modelEquationM4 = orig
    out["default_dataYs"] = np.array(res["dataYs"])
    out["default_series"] = series(res)
    out["default_X"] = np.array(res["XYList"][0][0])
    out["default_legs"] = np.array([d["leg"] for d in res["dataList"]])
    rng = np.random.default_rng(808)
    mi = cases.m8_case()
    y0 = np.array(list(mi["feed"]["concentration"]) + [523.0, 5e6, mi["feed"]["superficial-velocity"]])
    Y = np.tile(y0, (200, 1))
    Y[:, :6] *= np.exp(rng.normal(0, 0.4, (200, 6)))*10.0**(rng.uniform(-1, 2, (200, 6))*(y0[:6] < 1.0))
    Y[:, 6] = rng.uniform(490, 700, 200)
    Y[:, 7] *= rng.uniform(0.9, 1.0, 200)
    Y[:, 8] *= rng.uniform(0.7, 1.3, 200)
    Y[0] = y0
    with ref.quiet():
        out["rhs_y"], out["rhs_f"] = Y, np.array([orig(0.5, list(y), *cap["args"]) for y in Y])
    t = time.time()
    res = run(cases.m8_case(), rtol=1e-12, atol=1e-12)
    out["conv_series"] = series(res)
    rows = cases.ensemble_rows(8)
    out["sweep_conv_series"] = np.array([series(run(cases.m8_case(T=float(rows["T"][i]), P=float(rows["P"][i]), conc=rows["conc"][i],
                                                                 Tm=float(rows["T"][i])), rtol=1e-12, atol=1e-12)) for i in range(8)])
    print(f"m8 converged runs {time.time()-t:.1f}s", flush=True)
    save("m8", **out)


def g_m9():
    """Dynamic dimensional model M9 (rmtCore.py:322-331 -> pbReactor.py runM5 / modelEquationM5): RHS on random states at two
    grids, and rmtExe (which reports the temperature profile of every interval end) at reduced grids -- one reference solve at the
    default grid (100 nodes, Python loops, a print per RHS call) takes hours."""
    import functools
    import scipy.integrate as si
    PyREMOT, H, setting = ref.load()
    import PyREMOT.docs.pbReactor as R
    cls = R.PackedBedReactorClass
    orig = cls.modelEquationM5
    cap = {}

    def spy(t, y, a, b, c):
        cap["args"] = (a, b, c)
        raise RuntimeError("captured")

    def run(mi, zNo, tNo, rtol=None, atol=None):
        saved, old = R.solve_ivp, dict(setting["S2"])
        setting["S2"].update({"zNo": zNo, "tNo": tNo, "timesNo": 5})
        try:
            if rtol is not None:
                R.solve_ivp = functools.partial(si.solve_ivp, rtol=rtol, atol=atol)
            with ref.quiet():
                return PyREMOT.rmtExe(mi)["resModel"]
        finally:
            R.solve_ivp = saved
            setting["S2"].update(old)
    out = {}
    rng = np.random.default_rng(909)
    for zNo in (12, 40):
        cls.modelEquationM5 = spy
        try:
            run(cases.m9_case(), zNo, 2)
        except Exception:
            pass
        finally:
            cls.modelEquationM5 = orig
        mi = cases.m9_case()
        K = 12 if zNo == 12 else 4
        Y = np.zeros((K, 7, zNo))
        c0 = np.array(mi["feed"]["concentration"])
        Y[:, :6, :] = c0[None, :, None]*np.exp(rng.normal(0, 0.3, (K, 6, zNo)))*10.0**(rng.uniform(-1, 2, (K, 6, zNo))*(c0[None, :, None] < 1.0))
        Y[:, 6, :] = rng.uniform(500, 650, (K, zNo))
        Y[1::3, 2, zNo//2] = -1e-7               # below the 1e-30 clamp (:2424, :2552)
        with ref.quiet():
            F = np.array([np.array(orig(0.5, y.reshape(-1), *cap["args"])).reshape(7, zNo) for y in Y])
        out[f"rhs_y_z{zNo}"], out[f"rhs_f_z{zNo}"] = Y, F
    t = time.time()
    for tag, mi, zNo, tNo in (("z12", cases.m9_case(period=0.2), 12, 2), ("z20", cases.m9_case(period=0.1, T=540.0, Tm=530.0), 20, 2)):
        res = run(mi, zNo, tNo)
        out[f"{tag}_default_T"] = np.array([xy[1] for xy in res["XYList"]])
        out[f"{tag}_legs"] = np.array([d["leg"] for d in res["dataList"]])
        out[f"{tag}_X"] = np.array(res["XYList"][0][0])
        res = run(mi, zNo, tNo, rtol=1e-10, atol=1e-12)
        out[f"{tag}_conv_T"] = np.array([xy[1] for xy in res["XYList"]])
        print(f"m9 {tag} {time.time()-t:.1f}s", flush=True)
    save("m9", **out)


GROUPS = {"m7": g_m7, "m8": g_m8, "m9": g_m9, "props": g_props, "n1_rhs": g_n1_rhs, "n2_rhs": g_n2_rhs, "n1_solve": g_n1_solve,
          "n1_sweep": g_n1_sweep, "n2_solve": g_n2_solve}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--slow", action="store_true")
    a = ap.parse_args()
    names = a.only.split(",") if a.only else [g for g in GROUPS if a.slow or g not in ("n2_solve",)]
    for g in names:
        t = time.time()
        GROUPS[g]()
        print(f"[{g}] {time.time()-t:.1f}s", flush=True)
