"""Pin the CPU oracle (oracle/pbr_oracle.py) against outputs of the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from /root/reference
(PyREMOT 1.0.17, SciPy 1.18.1). RHS-level tolerance 1e-12 relative (pow/T**3 last-bit differences),
converged solves 1e-8.
"""
import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)/(np.abs(b) + floor))


def test_props_known_answers(golden):
    g = golden("props")
    sym = list(g["symbols"])
    assert tuple(sym) == O.COMPONENT_SYMBOLS                      # rmtCom() order
    Ts = g["Ts"]
    assert relerr(O.cp(sym, Ts), g["cp_all"]) < 1e-14
    assert relerr(O.cp_mean(sym, Ts), g["cpmean_all"]) < 1e-14
    assert relerr(O.viscosity(sym, Ts), g["mu_all"]) < 1e-14
    comps = cases.MEOH_DME_COMPONENTS
    S, dH25, sides = O.parse_reactions(cases.MEOH_DME_REACTIONS, comps)
    assert np.array_equal(S, [[-3, -1, 1, 0, 1, 0], [1, 1, -1, -1, 0, 0], [0, 0, 1, 0, -2, 1]])
    assert relerr(dH25, g["dH25"]) < 1e-14
    assert relerr(O.parse_reactions({"R1": "2CH4 <=> C2H4 + 2H2"}, ["CH4", "C2H4", "H2"])[1], g["dH25_ch4"]) < 1e-14
    assert relerr(O.dH_correction(sides, Ts), g["dHcorr"], floor=1e-9) < 1e-13
    MW = np.array([O._DB[s][0] for s in comps])
    mu = O.viscosity(comps, Ts)
    assert relerr(O.wilke(mu, np.broadcast_to(g["x0"], mu.shape), MW), g["mu_mix"]) < 1e-14
    # values printed by the reference's own scripts (tests/test_heatCapacity.py, test_viscosity.py)
    assert np.allclose(dH25, [-49010, -41160, -24520], rtol=1e-12)
    assert np.allclose(O.cp_mean(comps, 523.0), [28.605201550287504, 41.085448574875535, 34.750290853275004,
                                                 29.62105503875, 53.714055370462724, 79.34119544325], rtol=1e-14)
    assert np.allclose(O.wilke(O.viscosity(comps, 523.0), g["x0"], MW), 2.5105175291507334e-05, rtol=1e-13)


@pytest.mark.parametrize("tag,pt", [("noniso", "non-iso-thermal"), ("iso", "iso-thermal")])
def test_n1_rhs_vs_reference(golden, tag, pt):
    g = golden("n1_rhs")
    Y, F = g[f"y_{tag}"], g[f"f_{tag}"]
    ens = O.Ensemble(cases.readme_case("N1", pt), N=len(Y))
    assert np.allclose(ens.iv_n1()[0], Y[0], rtol=1e-15)
    f = ens.rhs_n1(Y)
    assert relerr(f, F, floor=1e-300) < 1e-11
    # scalar (non-vectorised) rate evaluation agrees with the numpy rebinding
    e1 = O.Ensemble(cases.readme_case("N1", pt), N=4)
    r_v = O.rates(e1.rate_expr, [523.0]*4, [5e6]*4, e1.x0, e1.C0)
    r_s = O.rates(e1.rate_expr, [523.0]*4, [5e6]*4, e1.x0, e1.C0, vectorised=False)
    assert relerr(r_v, r_s) < 1e-13
    if tag == "noniso":  # SURVEY.md §8c inlet values
        assert np.allclose(r_s[0], [85.04367720386661, -159.23189212928682, 236.7358718211693], rtol=1e-10)
        assert np.allclose(f[0], [-3.6040866831177381, -2.1246841259261795, 4.183788727621879, 1.3849828473304004,
                                  -3.3785079247956205, 2.0591046016956995, -1.8090067614742154e-03,
                                  2.6597909560060318], rtol=1e-10)


def test_n1_rhs_sweep_vs_reference(golden):
    g = golden("n1_rhs")
    rows = cases.ensemble_rows(24)
    ens = O.Ensemble(cases.readme_case("N1"), {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})
    for k in range(2):
        assert relerr(ens.rhs_n1(g["y_sweep"][:, k]), g["f_sweep"][:, k], floor=1e-300) < 1e-11
    assert np.allclose(ens.iv_n1(), g["y_sweep"][:, 0], rtol=1e-15)


@pytest.mark.parametrize("tag,mk,zNo", [
    ("readme_z5", lambda: cases.readme_case("N2"), 5),
    ("readme_z12", lambda: cases.readme_case("N2"), 12),
    ("readme_iso_z8", lambda: cases.readme_case("N2", "iso-thermal"), 8),
    ("dme_z20", cases.dme_n2_case, 20),
    ("ch4_z10", cases.ch4_n2_case, 10),
])
def test_n2_rhs_vs_reference(golden, tag, mk, zNo):
    g = golden("n2_rhs")
    Y, F = g[f"y_{tag}"], g[f"f_{tag}"]
    ens = O.Ensemble(mk(), N=len(Y))
    assert np.allclose(ens.iv_n2(zNo)[0], Y[0], rtol=1e-15)
    f = ens.rhs_n2(Y)
    assert f.shape == F.shape
    assert relerr(f, F, floor=1e-9) < 1e-11
    if tag == "readme_z5":
        assert np.allclose(f[0][:, 0], [-1.8481007150019408, -1.0894938433835588, 2.1453598702699304,
                                        0.7101904075743676, -1.7324286179635533, 1.0558660268863722,
                                        0.5319153449678685], rtol=1e-10)


def test_m7_oracle_vs_reference(golden):
    """Dimensional steady-state model M7: the oracle's restatement of modelEquationM3 (`pbReactor.py:1371-1577`) against 200
    reference states, and its SciPy solve against the reference's converged rmtExe("M7") profile and 2 swept rows."""
    g = golden("m7")
    ens = O.Ensemble(cases.m7_case(), N=len(g["rhs_y"]))
    f = ens.rhs_m7(g["rhs_y"])
    assert np.max(np.abs(f - g["rhs_f"])/(np.abs(g["rhs_f"]) + 1e-300)) < 1e-12
    _, data, _ = O.solve_m7(O.Ensemble(cases.m7_case()), method="LSODA", rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(data[0] - g["conv_dataYs"])/np.abs(g["conv_dataYs"])) < 1e-8
    rows = cases.ensemble_rows(8)
    sw = {"T": rows["T"][:2], "P": rows["P"][:2], "conc": rows["conc"][:2], "Tm": rows["T"][:2]}
    _, data, _ = O.solve_m7(O.Ensemble(cases.m7_case(), sw), method="LSODA", rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(data - g["sweep_conv_dataYs"][:2])/np.abs(g["sweep_conv_dataYs"][:2])) < 1e-8


def test_m8_oracle_vs_reference(golden):
    """Model M8: the oracle's restatement of modelEquationM4 (`pbReactor.py:1778-1995`, built-in kinetics `:8850-8930`) against 200
    reference states, and its SciPy solve against the reference's converged rmtExe("M8") series."""
    g = golden("m8")
    ens = O.Ensemble(cases.m8_case(), N=len(g["rhs_y"]))
    f = ens.rhs_m8(g["rhs_y"])
    assert np.max(np.abs(f - g["rhs_f"])/(np.abs(g["rhs_f"]) + 1e-300)) < 1e-12
    _, series, _ = O.solve_m8(O.Ensemble(cases.m8_case()), rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(series[0] - g["conv_series"])/np.abs(g["conv_series"])) < 1e-8


def test_m9_oracle_vs_reference(golden):
    """Model M9: the oracle's restatement of modelEquationM5 (`pbReactor.py:2296-2666`: marched pressure AND velocity, the clamps,
    the catalyst's heat capacity) against the reference's RHS on random states at two grids (one species below the 1e-30 clamp), and
    its interval loop against the reference's rmtExe("M9") temperature profiles -- converged and at the reference's tolerances."""
    g = golden("m9")
    for zNo in (12, 40):
        Y, F = g[f"rhs_y_z{zNo}"], g[f"rhs_f_z{zNo}"]
        f = O.Ensemble(cases.m9_case(), N=len(Y)).rhs_m9(Y)
        assert np.max(np.abs(f - F)/(np.abs(F) + 1e-12*np.abs(F).max())) < 1e-11, zNo
    for tag, mi, zNo in (("z12", cases.m9_case(period=0.2), 12),):     # (the z20 fixture is checked through the device code on the host)
        _, T, _ = O.solve_m9(O.Ensemble(mi), zNo=zNo, tNo=2, method="BDF", rtol=1e-10, atol=1e-12, sparse=True)
        assert np.max(np.abs(T[0] - g[f"{tag}_conv_T"]))/500 < 1e-10, tag
        _, T, _ = O.solve_m9(O.Ensemble(mi), zNo=zNo, tNo=2)                       # LSODA, rtol 1e-3: the reference's own run
        assert np.max(np.abs(T[0] - g[f"{tag}_default_T"]))/500 < 1e-9, tag
