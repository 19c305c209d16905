"""User-level example inputs (``modelInput`` dicts) and the synthetic ensembles of SURVEY.md §8(d).

These are *inputs*, in exactly the schema the reference's ``rmtExe`` reads
(`PyREMOT/docs/pbHomoReactor.py:2708-2782`, `:3332-3416`); nothing here computes.

* :func:`readme_case` – the methanol/DME example of `README.md:61-228`
  (kinetics `README.md:125-157, 169-173`).
* :func:`dme_n2_case` – the only runnable N-model script of the reference,
  `PyREMOT/tests/test_rmt_N1_DME.py` (runs model "N2").
* :func:`ch4_n2_case` – second kinetics fixture, `PyREMOT/tests/test_rmt_N2_CH4.py:188-202`.
* :func:`ensemble_rows` – the seeded synthetic sweeps (BASELINE.json configs 2-5).
"""
from __future__ import annotations

import math

import numpy as np

R_CONST = 8.314472  # PyREMOT/core/constants.py:8

MEOH_DME_COMPONENTS = ["H2", "CO2", "H2O", "CO", "CH3OH", "DME"]
MEOH_DME_REACTIONS = {
    "R1": "CO2 + 3H2 <=> CH3OH + H2O",
    "R2": "CO + H2O <=> H2 + CO2",
    "R3": "2CH3OH <=> DME + H2O",
}

# The six kinetic constants config 5 perturbs:
# pre-exponentials of K1..K3 and of the adsorption constants KH2, KCO2, KCO.
MEOH_DME_THETA = {
    "A_K1": 35.45, "A_K2": 7.3976, "A_K3": 8.2894e4,
    "A_KH2": 0.249, "A_KCO2": 1.02e-7, "A_KCO": 7.99e-7,
}


def methanol_dme_kinetics(CaBeDe=1171.2, theta_as_vars=False):
    """VARS / RATES lambda dicts of the README methanol/DME kinetics.

    ``theta_as_vars=True`` promotes the six pre-exponential constants to scalar
    ``VARS`` entries so that they become per-instance
    runtime parameters; the arithmetic is unchanged.
    """
    v = {"CaBeDe": CaBeDe}
    if theta_as_vars:
        v.update(MEOH_DME_THETA)
        a = {k: (lambda x, k=k: x[k]) for k in MEOH_DME_THETA}
    else:
        a = {k: (lambda x, c=c: c) for k, c in MEOH_DME_THETA.items()}
    v.update({
        "RT": lambda x: x['R_CONST']*x['T'],
        "K1": lambda x: a["A_K1"](x)*math.exp(-1.7069e4/x['RT']),
        "K2": lambda x: a["A_K2"](x)*math.exp(-2.0436e4/x['RT']),
        "K3": lambda x: a["A_K3"](x)*math.exp(-5.2940e4/x['RT']),
        "KH2": lambda x: a["A_KH2"](x)*math.exp(3.4394e4/x['RT']),
        "KCO2": lambda x: a["A_KCO2"](x)*math.exp(6.74e4/x['RT']),
        "KCO": lambda x: a["A_KCO"](x)*math.exp(5.81e4/x['RT']),
        "Ln_KP1": lambda x: 4213/x['T'] - 5.752 * math.log(x['T']) - 1.707e-3*x['T'] + 2.682e-6 * (math.pow(x['T'], 2)) - 7.232e-10*(math.pow(x['T'], 3)) + 17.6,
        "KP1": lambda x: math.exp(x['Ln_KP1']),
        "log_KP2": lambda x: 2167/x['T'] - 0.5194 * math.log10(x['T']) + 1.037e-3*x['T'] - 2.331e-7 * (math.pow(x['T'], 2)) - 1.2777,
        "KP2": lambda x: math.pow(10, x['log_KP2']),
        "Ln_KP3": lambda x: 4019/x['T'] + 3.707 * math.log(x['T']) - 2.783e-3*x['T'] + 3.8e-7 * (math.pow(x['T'], 2)) - 6.56e-4/(math.pow(x['T'], 3)) - 26.64,
        "KP3": lambda x: math.exp(x['Ln_KP3']),
        "yi_H2": lambda x: x['MoFri'][0],
        "yi_CO2": lambda x: x['MoFri'][1],
        "yi_H2O": lambda x: x['MoFri'][2],
        "yi_CO": lambda x: x['MoFri'][3],
        "yi_CH3OH": lambda x: x['MoFri'][4],
        "yi_DME": lambda x: x['MoFri'][5],
        "PH2": lambda x: x['P']*(x['yi_H2'])*1e-5,
        "PCO2": lambda x: x['P']*(x['yi_CO2'])*1e-5,
        "PH2O": lambda x: x['P']*(x['yi_H2O'])*1e-5,
        "PCO": lambda x: x['P']*(x['yi_CO'])*1e-5,
        "PCH3OH": lambda x: x['P']*(x['yi_CH3OH'])*1e-5,
        "PCH3OCH3": lambda x: x['P']*(x['yi_DME'])*1e-5,
        "ra1": lambda x: x['PCO2']*x['PH2'],
        "ra2": lambda x: 1 + (x['KCO2']*x['PCO2']) + (x['KCO']*x['PCO']) + math.sqrt(x['KH2']*x['PH2']),
        "ra3": lambda x: (1/x['KP1'])*((x['PH2O']*x['PCH3OH'])/(x['PCO2']*(math.pow(x['PH2'], 3)))),
        "ra4": lambda x: x['PH2O'] - (1/x['KP2'])*((x['PCO2']*x['PH2'])/x['PCO']),
        "ra5": lambda x: (math.pow(x['PCH3OH'], 2)/x['PH2O'])-(x['PCH3OCH3']/x['KP3']),
    })
    r = {
        "r1": lambda x: 1000*x['K1']*(x['ra1']/(math.pow(x['ra2'], 3)))*(1-x['ra3'])*x['CaBeDe'],
        "r2": lambda x: 1000*x['K2']*(1/x['ra2'])*x['ra4']*x['CaBeDe'],
        "r3": lambda x: 1000*x['K3']*x['ra5']*x['CaBeDe'],
    }
    return v, r


# Dashboard text form of the same kinetics (`README.md:88-120, 160-164`, with the
# markdown mangling at :102-115 undone). Used to test the string front-end.
MEOH_DME_VARS_TEXT = """
"CaBeDe" : 1171.2;
"RT": x['R_CONST']*x['T'];
"K1": 35.45*math.exp(-1.7069e4/x['RT']);
"K2": 7.3976*math.exp(-2.0436e4/x['RT']);
"K3": 8.2894e4*math.exp(-5.2940e4/x['RT']);
"KH2": 0.249*math.exp(3.4394e4/x['RT']);
"KCO2": 1.02e-7*math.exp(6.74e4/x['RT']);
"KCO": 7.99e-7*math.exp(5.81e4/x['RT']);
"Ln_KP1": 4213/x['T'] - 5.752 * math.log(x['T']) - 1.707e-3*x['T'] + 2.682e-6 * (math.pow(x['T'], 2)) - 7.232e-10*(math.pow(x['T'], 3)) + 17.6;
"KP1": math.exp(x['Ln_KP1']);
"log_KP2": 2167/x['T'] - 0.5194 * math.log10(x['T']) + 1.037e-3*x['T'] - 2.331e-7 * (math.pow(x['T'], 2)) - 1.2777;
"KP2": math.pow(10, x['log_KP2']);
"Ln_KP3": 4019/x['T'] + 3.707 * math.log(x['T']) - 2.783e-3*x['T'] + 3.8e-7 * (math.pow(x['T'], 2)) - 6.56e-4/(math.pow(x['T'], 3)) - 26.64;
"KP3": math.exp(x['Ln_KP3']);
"yi_H2": x['MoFri'][0];
"yi_CO2": x['MoFri'][1];
"yi_H2O": x['MoFri'][2];
"yi_CO": x['MoFri'][3];
"yi_CH3OH": x['MoFri'][4];
"yi_DME": x['MoFri'][5];
"PH2": x['P']*(x['yi_H2'])*1e-5;
"PCO2": x['P']*(x['yi_CO2'])*1e-5;
"PH2O": x['P']*(x['yi_H2O'])*1e-5;
"PCO": x['P']*(x['yi_CO'])*1e-5;
"PCH3OH": x['P']*(x['yi_CH3OH'])*1e-5;
"PCH3OCH3": x['P']*(x['yi_DME'])*1e-5;
"ra1": x['PCO2']*x['PH2'];
"ra2": 1 + (x['KCO2']*x['PCO2']) + (x['KCO']*x['PCO']) + math.sqrt(x['KH2']*x['PH2']);
"ra3": (1/x['KP1'])*((x['PH2O']*x['PCH3OH'])/(x['PCO2']*(math.pow(x['PH2'], 3))));
"ra4": x['PH2O'] - (1/x['KP2'])*((x['PCO2']*x['PH2'])/x['PCO']);
"ra5": (math.pow(x['PCH3OH'], 2)/x['PH2O'])-(x['PCH3OCH3']/x['KP3'])
"""
MEOH_DME_RATES_TEXT = """
"r1": 1000*x['K1']*(x['ra1']/(math.pow(x['ra2'], 3)))*(1-x['ra3'])*x['CaBeDe'];
"r2": 1000*x['K2']*(1/x['ra2'])*x['ra4']*x['CaBeDe'];
"r3": 1000*x['K3']*x['ra5']*x['CaBeDe']
"""


def _model_input(model, T, P, period, process_type, VoFlRa, conc, comps, reactions, vars_, rates,
                 U, Tm, ReInDi, ReLe, PaDi, BeVoFr, CaBeDe, CaDe, CaSpHeCa, ivp):
    op = {"pressure": P, "temperature": T, "process-type": process_type}
    if period is not None:
        op["period"] = period
    return {
        "model": model,
        "operating-conditions": op,
        "feed": {
            "volumetric-flowrate": VoFlRa,
            "concentration": list(conc),
            "components": {"shell": list(comps)},
        },
        "reactions": dict(reactions),
        "reaction-rates": {"VARS": vars_, "RATES": rates},
        "external-heat": {"OvHeTrCo": U, "EfHeTrAr": 4/ReInDi, "MeTe": Tm},
        "reactor": {"ReInDi": ReInDi, "ReLe": ReLe, "PaDi": PaDi, "BeVoFr": BeVoFr,
                    "CaBeDe": CaBeDe, "CaDe": CaDe, "CaSpHeCa": CaSpHeCa},
        "solver-config": {"ivp": ivp, "display-result": "False"},
    }


def readme_case(model="N1", process_type="non-iso-thermal", ivp="default", period=0.5,
                T=523.0, P=5e6, conc=None, VoFlRa=0.000228, U=50.0, Tm=523.0,
                theta_as_vars=False):
    """`README.md:180-228` inputs; the N1 golden case of SURVEY.md §8(c)."""
    CaBeDe = 1171.2
    v, r = methanol_dme_kinetics(CaBeDe, theta_as_vars=theta_as_vars)
    if conc is None:
        conc = [574.8978, 287.4489, 1.15e-02, 287.4489, 1.15e-02, 1.15e-02]
    return _model_input(model, T, P, period if model == "N2" else None, process_type, VoFlRa, conc,
                        MEOH_DME_COMPONENTS, MEOH_DME_REACTIONS, v, r, U, Tm,
                        ReInDi=0.0381, ReLe=1, PaDi=0.002, BeVoFr=0.39, CaBeDe=CaBeDe,
                        CaDe=1920, CaSpHeCa=960, ivp=ivp)


def m7_case(ivp="default", T=523.0, P=5e6, conc=None, U=50.0, Tm=523.0, mixture_viscosity=2.5105175291507334e-05):
    """The README inputs run through the dimensional steady-state model "M7" (`PyREMOT/docs/rmtCore.py:300-309` ->
    `pbReactor.py` runM3 `:1170`): same schema plus ``feed['mixture-viscosity']`` [Pa.s] (`:1238`; here the Wilke value of the
    README feed) -- and ``external-heat['EfHeTrAr']`` IS used by this model (`:1545`)."""
    mi = readme_case("N1", ivp=ivp, T=T, P=P, conc=conc, U=U, Tm=Tm)
    mi["model"] = "M7"
    mi["feed"]["mixture-viscosity"] = mixture_viscosity
    return mi


def m8_case(T=523.0, P=5e6, conc=None, U=50.0, Tm=523.0, superficial_velocity=None):
    """The README inputs through the dimensional steady-state model "M8" (`PyREMOT/docs/rmtCore.py:311-320` -> `pbReactor.py`
    runM4 `:1580`): unknowns C_i, T, P and the superficial velocity v, whose inlet value is the input
    ``feed['superficial-velocity']`` (`:1626`; default here: volumetric flow rate / cross-section).  The reference integrates this
    model with its BUILT-IN methanol/DME kinetics (`modelReactions`, `:8850`), not with ``reaction-rates``."""
    mi = m7_case(T=T, P=P, conc=conc, U=U, Tm=Tm)
    mi["model"] = "M8"
    A = math.pi*mi["reactor"]["ReInDi"]**2/4
    mi["feed"]["superficial-velocity"] = mi["feed"]["volumetric-flowrate"]/A if superficial_velocity is None else superficial_velocity
    return mi


def m9_case(period=0.5, T=523.0, P=5e6, conc=None, U=50.0, Tm=523.0, ivp="default"):
    """The README inputs through the dynamic dimensional model "M9" (`PyREMOT/docs/rmtCore.py:322-331` -> `pbReactor.py` runM5
    `:1997`): C_i and T per node, pressure and velocity marched along the bed; same extra inputs as M8.  Grid and intervals come
    from ``solverSetting['S2']`` (`solSetting.py`: zNo 100, tNo 10)."""
    mi = m8_case(T=T, P=P, conc=conc, U=U, Tm=Tm)
    mi["model"] = "M9"
    mi["operating-conditions"]["period"] = period
    mi["solver-config"]["ivp"] = ivp
    return mi


def feed_mole_fraction(H2COxRatio, CO2COxRatio, float32=False):
    """Feed composition recipe of `PyREMOT/data/initData.py:11-38`.

    The reference helper casts to float32 (`:37-38`); ``float32=True`` reproduces
    that for the `dme_n2_case` fixture, the synthetic ensembles stay in float64.
    """
    H2COxRatio = np.asarray(H2COxRatio, dtype=np.float64)
    CO2COxRatio = np.asarray(CO2COxRatio, dtype=np.float64)
    tr = 0.00001
    tmf0 = 1 - (tr + tr + tr)
    COx = tmf0/(H2COxRatio + 1)
    y_H2 = H2COxRatio*COx
    y_CO2 = CO2COxRatio*COx
    y_CO = COx - y_CO2
    t = np.full_like(COx, tr)
    y = np.stack([y_H2, y_CO2, t, y_CO, t, t], axis=-1)
    return y.astype(np.float32) if float32 else y


def dme_n2_case(ivp="default", period=0.5):
    """Inputs of `PyREMOT/tests/test_rmt_N1_DME.py` (model "N2", U=100, Tm=T-1)."""
    P, T = 5e6, 523
    bed_por, rea_D, rea_L, cat_d, cat_rho, cat_Cp = 0.39, 0.0381, 1, 0.002, 1982, 960
    y = feed_mole_fraction(1, 0.5, float32=True)
    # calConcentration, `PyREMOT/data/initData.py:43-69`: kmol/m^3 rounded to 7 decimals
    # (CONCENTRATION_ACCURACY, `PyREMOT/core/config.py`), then *1e3 in the script (:49).
    # Under numpy>=2 the python-float factor is weak, so the product with the float32
    # mole fraction (and the /1000) is evaluated in float32 by the reference script.
    ct0 = np.round(np.array([np.float32(P/(R_CONST*T))*yi/np.float32(1000) for yi in y],
                            dtype=np.float64), 7)
    conc = 1e3*ct0
    InGaVe = 0.2/bed_por
    rea_CSA = bed_por*math.pi*rea_D**2/4  # rmtUtility.reactorCrossSectionArea
    VoFlRa = InGaVe*rea_CSA
    CaBeDe = cat_rho*(1-bed_por)
    v, r = methanol_dme_kinetics(CaBeDe)
    return _model_input("N2", T, P, period, "non-iso-thermal", VoFlRa, conc,
                        MEOH_DME_COMPONENTS, MEOH_DME_REACTIONS, v, r, U=100, Tm=T-1,
                        ReInDi=rea_D, ReLe=rea_L, PaDi=cat_d, BeVoFr=bed_por, CaBeDe=CaBeDe,
                        CaDe=cat_rho, CaSpHeCa=cat_Cp/1000, ivp=ivp)


def ch4_n2_case(ivp="default", period=10):
    """Inputs of `PyREMOT/tests/test_rmt_N2_CH4.py` (concentration-based rate, adiabatic Tm=0)."""
    P, T = 3e5, 973
    bed_por, rea_dia, cat_d, cat_rho, cat_cp = 0.39, 0.007, 0.002, 1982, 960
    y = np.array([0.9, 0.05, 0.05])
    ct0 = np.round((P/(R_CONST*T))*y/1000, 7)
    conc = 1e3*ct0
    InGaVe = 0.01/bed_por
    VoFlRa = InGaVe*bed_por*math.pi*rea_dia**2/4
    v = {
        "k0": 0.0072*1e-1,
        "y_CH4": lambda x: x['MoFri'][0],
        "C_CH4": lambda x: x['SpCoi'][0],
    }
    r = {"r1": lambda x: x['k0']*(x['C_CH4']**2)}
    return _model_input("N2", T, P, period, "non-iso-thermal", VoFlRa, conc,
                        ["CH4", "C2H4", "H2"], {"R1": "2CH4 <=> C2H4 + 2H2"}, v, r, U=50, Tm=0,
                        ReInDi=rea_dia, ReLe=1, PaDi=cat_d, BeVoFr=bed_por,
                        CaBeDe=cat_rho*(1-bed_por), CaDe=cat_rho, CaSpHeCa=cat_cp/1000, ivp=ivp)


ENSEMBLE_SEED = 20261019


def ensemble_rows(N, seed=ENSEMBLE_SEED, offset=0):
    """Synthetic sweep of SURVEY.md §8(d) item 2: rows ``offset .. offset+N`` of the seeded table.

    Returns a dict of per-instance arrays: ``T`` [K], ``P`` [Pa], ``conc`` [N,6] mol/m^3.
    ``T~U[493,553]``, ``P~U[3e6,7e6]``, ``H2/COx~U[1,3]``, ``CO2/COx~U[0.2,0.8]``, traces 1e-5,
    ``C0 = x0 P/(R T)``. Row i depends only on (seed, i): every rank regenerates its own shard.
    """
    rng = np.random.default_rng(seed)
    u = rng.random((offset + N, 4))[offset:]
    T = 493.0 + 60.0*u[:, 0]
    P = 3e6 + 4e6*u[:, 1]
    h2 = 1.0 + 2.0*u[:, 2]
    co2 = 0.2 + 0.6*u[:, 3]
    x0 = feed_mole_fraction(h2, co2)
    conc = x0*(P/(R_CONST*T))[:, None]
    return {"T": T, "P": P, "conc": conc, "H2COx": h2, "CO2COx": co2}
