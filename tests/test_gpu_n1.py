"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C-ABI (libpbr.so).

Tolerances: RHS / prolog values 1e-11 relative against reference-generated golden vectors (pow / T**3
last-bit differences); solves 1e-6 relative (the bar BASELINE.json's north_star states) on mole fractions,
P, T over the whole axial profile against the reference integrated at converged tolerances.
"""
import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases

pytestmark = pytest.mark.gpu

RTOL_SOLVE = 1e-6      # north_star tolerance
RTOL_RHS = 1e-11


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)/(np.abs(b) + floor)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


def _model(mi, params=()):
    from pyremot_b200.api import get_model
    return get_model(mi, params=params, device=0)


def _sweep24():
    rows = cases.ensemble_rows(24)
    return {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}


def test_prolog_matches_oracle(torch_cuda):
    torch = torch_cuda
    mi = cases.readme_case("N1")
    m = _model(mi)
    sweep = _sweep24()
    par = m.prolog(torch.from_numpy(m.instance_table(mi, sweep)).cuda()).cpu().numpy()
    ens = O.Ensemble(mi, sweep)
    eps, dp = ens.BeVoFr, ens.PaDi
    ref = {0: ens.T0, 1: ens.P0, 2: ens.Cmax, 3: ens.Ct0, 4: ens.vs0, 5: eps, 6: ens.ReLe,
           7: 150*ens.mu*((1-eps)**2)/(eps**3)/(dp**2), 8: 1.75*(1-eps)/(eps**3)/dp, 9: ens.rho0,
           10: ens.Cpmix0, 11: ens.G, 12: ens.GH, 13: ens.U*ens.a, 14: ens.Tm, 15: ens.mu}
    for row, v in ref.items():
        assert relerr(par[row], v) < 1e-13, row
    assert relerr(par[16:22].T, ens.c0) < 1e-15


@pytest.mark.parametrize("tag,pt", [("noniso", "non-iso-thermal"), ("iso", "iso-thermal")])
def test_rhs_n1_vs_reference_golden(torch_cuda, golden, tag, pt):
    torch = torch_cuda
    g = golden("n1_rhs")
    Y, F = g[f"y_{tag}"], g[f"f_{tag}"]
    mi = cases.readme_case("N1", pt)
    m = _model(mi)
    par = m.prolog(torch.from_numpy(m.instance_table(mi, N=len(Y))).cuda())
    f = m.rhs_n1(par, torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()).cpu().numpy().T
    assert relerr(f, F, floor=1e-300) < RTOL_RHS
    f2, J = m.jac_n1(par, torch.from_numpy(np.ascontiguousarray(Y.T)).cuda())
    # same formulas through pbr_rates_jac instead of pbr_rates: identical up to the compiler's FMA contraction choices
    assert relerr(f2.cpu().numpy().T, f, floor=1e-300) < 1e-12


def test_rhs_n1_sweep_and_jacobian(torch_cuda, golden):
    torch = torch_cuda
    g = golden("n1_rhs")
    mi = cases.readme_case("N1")
    m = _model(mi)
    sweep = _sweep24()
    par = m.prolog(torch.from_numpy(m.instance_table(mi, sweep)).cuda())
    ens = O.Ensemble(mi, sweep)
    for k in range(2):
        y = np.ascontiguousarray(g["y_sweep"][:, k].T)
        f, J = m.jac_n1(par, torch.from_numpy(y).cuda())
        assert relerr(f.cpu().numpy().T, g["f_sweep"][:, k], floor=1e-300) < RTOL_RHS
        J = J.cpu().numpy()
        if k == 1:   # analytic Jacobian against central differences of the oracle RHS
            for i in (0, 7, 23):
                e = ens.sub(i)
                Jfd = O.fd_jacobian(lambda yy: e.rhs_n1(yy[None])[0], y[:, i], rel=1e-6)
                colmax = np.max(np.abs(Jfd), axis=0, keepdims=True)
                assert np.max(np.abs(J[:, :, i] - Jfd)/colmax) < 1e-5


def test_rate_parameters_as_runtime_inputs(torch_cuda):
    """Scalar VARS promoted to per-instance parameters (config 5 mechanism) give the same RHS as constants."""
    torch = torch_cuda
    mi = cases.readme_case("N1")
    mi_p = cases.readme_case("N1", theta_as_vars=True)
    m, mp = _model(mi), _model(mi_p, params=tuple(cases.MEOH_DME_THETA))
    sweep = _sweep24()
    y = torch.from_numpy(np.ascontiguousarray(O.Ensemble(mi, sweep).iv_n1().T)).cuda()
    f0 = m.rhs_n1(m.prolog(torch.from_numpy(m.instance_table(mi, sweep)).cuda()), y).cpu().numpy()
    fp = mp.rhs_n1(mp.prolog(torch.from_numpy(mp.instance_table(mi_p, sweep)).cuda()), y).cpu().numpy()
    assert relerr(fp, f0, floor=1e-300) < 1e-13
    sw2 = dict(sweep)
    sw2["VARS:A_K1"] = np.full(24, 35.45*1.5)
    f2 = mp.rhs_n1(mp.prolog(torch.from_numpy(mp.instance_table(mi_p, sw2)).cuda()), y).cpu().numpy()
    e2 = O.Ensemble(mi_p, sw2)
    assert relerr(f2.T, e2.rhs_n1(y.cpu().numpy().T), floor=1e-300) < RTOL_RHS
    assert relerr(f2, f0, floor=1e-300) > 1e-3


@pytest.mark.parametrize("tag,pt", [("noniso", "non-iso-thermal"), ("iso", "iso-thermal")])
def test_rmtExe_n1_matches_converged_reference(torch_cuda, golden, tag, pt):
    from pyremot_b200 import rmtExe
    g = golden("n1_solve")
    res = rmtExe(cases.readme_case("N1", pt))
    assert {"resModel", "comTime"} <= set(res) and res["solver"]["ivp_requested"] == "default" and res["solver"]["rtol"] == 1e-9
    d = res["resModel"][0]
    assert set(d) == {"modelId", "processType", "successStatus", "computation-time", "dataShape", "labelList",
                      "indexList", "dataTime", "dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2", "dataYs"}
    assert list(d["labelList"]) == list(g[f"{tag}_labelList"]) and list(d["indexList"]) == list(g[f"{tag}_indexList"])
    assert tuple(d["dataShape"]) == tuple(g[f"{tag}_dataShape"]) and d["dataTime"] == [] and d["successStatus"] is True
    for k in ("dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2", "dataYs"):
        ref = g[f"{tag}_conv_{k}"]
        assert np.shape(d[k]) == ref.shape, k
        assert relerr(d[k], ref, floor=1e-9 if k in ("dataXs", "dataYTemp1") else 0.0) < RTOL_SOLVE, k
    # distance to what the reference returns at its own default tolerances (it runs solve_ivp at rtol 1e-3)
    for ivp in ("default", "BDF", "Radau"):
        assert relerr(d["dataYs"][:, -1], g[f"{tag}_{ivp}_dataYs"][:, -1]) < 2e-3
    if tag == "noniso":   # SURVEY.md §8c converged outlet
        assert np.allclose(d["dataYs"][:, -1], [0.4529955649935, 0.2615213300567, 0.02404559612944, 0.2384624972004,
                                                0.006905891959625, 0.01606911966034, 4992662.119233, 620.8459970187],
                           rtol=1e-7)


def test_batch_sweep_matches_converged_reference(torch_cuda, golden):
    """16 corners + 16 random rows of the config-2 box, full 101-point profiles, device and host-buffer paths."""
    from pyremot_b200 import rmtExeBatch
    gs = golden("n1_sweep")
    mi = cases.readme_case("N1")
    sweep = {"T": gs["T"], "P": gs["P"], "conc": gs["conc"], "Tm": gs["T"]}
    out = rmtExeBatch(mi, sweep, profile=True)
    assert (out["status"] == 0).all()
    assert relerr(out["dataYs"], gs["dataYs"]) < RTOL_SOLVE
    assert relerr(out["outlet"], gs["outlet"]) < RTOL_SOLVE
    assert np.array_equal(out["outlet"], out["dataYs"][:, :, -1])
    dev = rmtExeBatch(mi, sweep, profile=True, return_device=True)
    assert np.array_equal(dev["dataYs"].cpu().numpy(), out["dataYs"])
    assert np.array_equal(dev["outlet"].cpu().numpy(), out["outlet"])


def test_error_behaviour(torch_cuda):
    from pyremot_b200 import rmtExe
    mi = cases.readme_case("N1")
    mi["feed"]["components"]["shell"] = ["H2", "CO2", "H2O", "CO", "CH3OH", "XYZ"]
    with pytest.raises(Exception, match="Component database is not up to date!"):
        rmtExe(mi)
    mi = cases.readme_case("N1")
    del mi["feed"]["volumetric-flowrate"]
    with pytest.raises(KeyError):
        rmtExe(mi)
    mi = cases.readme_case("N1")
    mi["reaction-rates"]["VARS"] = {"K1": lambda x: x["nope"]*2.0, **mi["reaction-rates"]["VARS"]}
    with pytest.raises(KeyError):       # forward reference, as in rmtReaction.py:44-51
        rmtExe(mi)


def test_full_size_ensemble_properties(torch_cuda):
    """BASELINE config 2 at full size (65,536 reactors): size-independent invariants.

    * every instance integrates (status 0);
    * mole fractions sum to 1;
    * element totals sum_i a_ik c_i are linear invariants of dc/dz = (..) S^T r (each reaction conserves C, H, O),
      and BDF preserves linear invariants: C:H:O ratios of the outlet mixture equal the feed's;
    * determinism: a second run is bit-identical; a sub-batch reproduces the same rows (no cross-instance coupling).
    """
    from pyremot_b200 import rmtExeBatch
    N = 65536
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N1")
    out = rmtExeBatch(mi, sweep)
    assert (out["status"] == 0).all()
    x = out["outlet"][:, :6]
    assert np.max(np.abs(x.sum(axis=1) - 1)) < 1e-12
    #           H2 CO2 H2O CO CH3OH DME
    A = np.array([[0, 1, 0, 1, 1, 2],      # C
                  [2, 0, 2, 0, 4, 6],      # H
                  [0, 2, 1, 1, 1, 1]])     # O
    x0 = rows["conc"]/rows["conc"].sum(axis=1, keepdims=True)
    e0, e1 = x0 @ A.T, x @ A.T
    assert relerr(e1[:, 1]/e1[:, 0], e0[:, 1]/e0[:, 0]) < 1e-9
    assert relerr(e1[:, 2]/e1[:, 0], e0[:, 2]/e0[:, 0]) < 1e-9
    T, P = out["outlet"][:, 7], out["outlet"][:, 6]
    assert T.min() > 493 and T.max() < 700 and (P < rows["P"]).all() and (P > 0.99*rows["P"]).all()
    out2 = rmtExeBatch(mi, sweep)
    assert np.array_equal(out2["outlet"], out["outlet"])
    sub = rmtExeBatch(mi, {k: v[1000:1100] for k, v in sweep.items()})
    assert np.array_equal(sub["outlet"], out["outlet"][1000:1100])
    # ---- parity where the margin is thin: 256 seeded rows + the 32 rows that needed the most RHS evaluations +
    # full 101-point profiles of 16 rows (8 hardest, 8 seeded), all against the oracle integrated at converged tolerance
    import parity_helpers as H
    rng = np.random.default_rng(7)
    seeded = sorted(rng.choice(N, 256, replace=False).tolist())
    hardest = np.argsort(out["stats"][2])[-32:].tolist()
    idx = sorted(set(seeded + hardest))
    ref = H.oracle_n1(sweep, idx)
    err = H.relerr_rows(out["outlet"][idx], ref)
    worst = int(np.argmax(err))
    print(f"\nconfig 2: {len(idx)} outlets vs oracle: worst rel. error {err.max():.2e} at row {idx[worst]} "
          f"(nfev {out['stats'][2][idx[worst]]}, T0 {rows['T'][idx[worst]]:.1f} K); hardest-32 worst {H.relerr_rows(out['outlet'][hardest], ref[[idx.index(i) for i in hardest]]).max():.2e}")
    assert err.max() < RTOL_SOLVE/3, (err.max(), idx[worst])
    pidx = hardest[-8:] + seeded[:8]
    prof = rmtExeBatch(mi, {k: v[pidx] for k, v in sweep.items()}, profile=True)
    assert np.array_equal(prof["outlet"], out["outlet"][pidx])
    pref = H.oracle_n1(sweep, pidx, profiles=True)
    perr = H.relerr_rows(prof["dataYs"], pref)
    print(f"config 2: 16 full profiles (101 points) vs oracle: worst rel. error {perr.max():.2e}")
    assert perr.max() < RTOL_SOLVE/3, perr


def test_box_corners_vs_oracle(torch_cuda):
    """The 16 corners of the config-2 box (T in {493, 553} K, P in {3, 7} MPa, H2/COx in {1, 3}, CO2/COx in {0.2, 0.8}):
    the extremes of stiffness and hot-spot height, full profiles against the oracle."""
    import itertools
    import parity_helpers as H
    from pyremot_b200 import rmtExeBatch
    c = np.array(list(itertools.product((493.0, 553.0), (3e6, 7e6), (1.0, 3.0), (0.2, 0.8))))
    x0 = cases.feed_mole_fraction(c[:, 2], c[:, 3])
    sweep = {"T": c[:, 0], "P": c[:, 1], "conc": x0*(c[:, 1]/(cases.R_CONST*c[:, 0]))[:, None], "Tm": c[:, 0]}
    out = rmtExeBatch(cases.readme_case("N1"), sweep, profile=True)
    assert (out["status"] == 0).all()
    ref = H.oracle_n1(sweep, range(16), profiles=True)
    err = H.relerr_rows(out["dataYs"], ref)
    print(f"\nbox corners: worst profile rel. error {err.max():.2e} (T_max {out['dataYs'][:, 7].max():.1f} K)")
    assert err.max() < RTOL_SOLVE/3, err


def test_rmtExe_with_dashboard_text_rates(torch_cuda, golden):
    """the dashboard's textual VARS / RATES (`README.md:88-120, 160-164`) straight through rmtExe on the GPU.
    The text compiles to the same program as the lambdas (same cubin), so the result must be bit-identical to the lambda run and
    within 1e-6 of the reference's converged profile."""
    from pyremot_b200 import rmtExe
    g = golden("n1_solve")
    mi = cases.readme_case("N1")
    lam = rmtExe(mi)["resModel"][0]
    mi_t = cases.readme_case("N1")
    mi_t["reaction-rates"] = {"VARS": cases.MEOH_DME_VARS_TEXT, "RATES": cases.MEOH_DME_RATES_TEXT}
    txt = rmtExe(mi_t)["resModel"][0]
    assert np.array_equal(txt["dataYs"], lam["dataYs"])
    assert relerr(txt["dataYs"], g["noniso_conv_dataYs"]) < RTOL_SOLVE


def test_empty_and_tiny_shards(torch_cuda):
    """N = 0 (a rank whose shard is empty) must be a no-op, not a crash (integer division by zero in the launcher)."""
    torch = torch_cuda
    mi = cases.readme_case("N1")
    m = _model(mi)
    tab = torch.empty((0, m.k_in), dtype=torch.float64, device="cuda")
    par = m.prolog(tab)
    r = m.solve_n1(par, profile=False)
    assert r["outlet_y"].shape == (m.nv1, 0) and r["status"].shape == (0,)
    assert m.unscale_n1(par, r["outlet_y"][None]).shape == (0, m.nv1, 1)
    m2 = _model(cases.readme_case("N2"))
    par2 = m2.prolog(torch.empty((0, m2.k_in), dtype=torch.float64, device="cuda"))
    assert m2.solve_n2(par2, 16, 2, 0.1)["snaps_y"].shape == (0, 2, m2.nv2, 16)


def test_parameter_estimation_fd_jacobian(torch_cuda):
    """BASELINE config 5 mechanism: d(outlet)/d(theta) for the six kinetic constants from one ensemble launch
    (N x 7 solves), against the oracle differentiated the same way (SciPy at converged tolerance)."""
    from pyremot_b200 import estimation
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    rows = cases.ensemble_rows(512)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    J, base = estimation.outlet_jacobian_fd(mi, sweep, theta, rel=1e-4)
    assert J.shape == (512, 8, 6) and np.isfinite(J).all()
    # all six columns on 8 base reactors against the oracle differentiated the same way (8 x 7 converged SciPy solves)
    import parity_helpers as H
    base_idx = [17, 3, 100, 200, 255, 300, 411, 511]
    nb = len(base_idx)
    sw = {k: np.concatenate([v[base_idx]]*7) for k, v in sweep.items()}
    hs = []
    for k, name in enumerate(theta):
        col = np.full(7*nb, cases.MEOH_DME_THETA[name])
        h = cases.MEOH_DME_THETA[name]*1e-4
        col[(k + 1)*nb:(k + 2)*nb] += h
        sw["VARS:" + name] = col
        hs.append(h)
    ref = H.oracle_n1(sw, range(7*nb), rtol=1e-12, atol=1e-15, theta_as_vars=True).reshape(7, nb, 8)
    assert relerr(base[base_idx], ref[0]) < RTOL_SOLVE
    worst = 0.0
    for k in range(6):
        Jref = (ref[k + 1] - ref[0])/hs[k]                            # [nb, 8]
        for b, i in enumerate(base_idx):
            scale = np.max(np.abs(Jref[b, :6]))
            # FD noise: 1e-11/1e-4 relative to the column's scale, plus the integration-noise floor of the two solves
            # (rtol 1e-11 here, 1e-12 in the oracle) divided by the step: it dominates where the outlet sits at the chemical
            # equilibrium of a reaction (row 17, K3: d(outlet)/dK3 ~ 6e-12, i.e. outlet differences of 5e-11)
            floor = 5e-12*np.max(np.abs(base[i, :6]))/hs[k]
            d = np.max(np.abs(J[i, :6, k] - Jref[b, :6]))
            worst = max(worst, d/(2e-3*scale + floor))
            assert d < 2e-3*scale + floor, (i, k, d, scale, floor)
    print(f"\nconfig 5: 8 reactors x 6 columns vs oracle FD: worst deviation {worst:.2f} of the allowed band")


def test_failures_are_reported_per_instance_not_hung(torch_cuda):
    """Broken inputs must terminate with a status code (bounded step attempts), and must not disturb their neighbours."""
    from pyremot_b200 import rmtExeBatch
    mi = cases.readme_case("N1")
    rows = cases.ensemble_rows(64)
    sweep = {"T": rows["T"].copy(), "P": rows["P"].copy(), "conc": rows["conc"].copy(), "Tm": rows["T"].copy()}
    good = rmtExeBatch(mi, sweep)
    sweep["T"][5] = float("nan")
    sweep["conc"][9] = 0.0                       # empty feed: Ct0 = 0 -> non-finite constants
    sweep["P"][33] = -5e6
    out = rmtExeBatch(mi, sweep)
    bad = {5, 9, 33}
    assert all(out["status"][i] != 0 for i in bad)
    assert np.isnan(out["outlet"][sorted(bad)]).all()          # failed rows are NaN, never plausible-looking numbers
    ok = [i for i in range(64) if i not in bad]
    assert (out["status"][ok] == 0).all()
    assert np.array_equal(out["outlet"][ok], good["outlet"][ok])
    with pytest.raises(RuntimeError, match="3 of 64"):
        rmtExeBatch(mi, sweep, strict=True)
    # N = 1 and an N that is not a multiple of the block size
    one = rmtExeBatch(mi, {k: v[:1] for k, v in sweep.items()})
    assert np.array_equal(one["outlet"], good["outlet"][:1])
    odd = rmtExeBatch(mi, {k: v[10:33] for k, v in sweep.items()})
    assert np.array_equal(odd["outlet"], good["outlet"][10:33])


def test_other_kinetics_and_process_types_vs_oracle(torch_cuda):
    """Model shapes other than the methanol/DME non-iso-thermal one: the CH4 coupling kinetics of
    PyREMOT/tests/test_rmt_N2_CH4.py run as N1 (3 species, 1 reaction, SpCoi-based rate, adiabatic Tm = 0: nv1 = 5,
    256-thread blocks) and an iso-thermal methanol/DME sweep (nv1 = 7), both against the oracle at converged tolerance."""
    from pyremot_b200 import rmtExeBatch
    mi = cases.ch4_n2_case()
    mi["model"] = "N1"
    N = 300                                             # not a multiple of the block size
    rng = np.random.default_rng(5)
    sweep = {"T": 900 + 150*rng.random(N), "P": 2e5 + 2e5*rng.random(N)}
    sweep["conc"] = np.array([0.9, 0.05, 0.05])[None, :]*(sweep["P"]/(cases.R_CONST*sweep["T"]))[:, None]
    # 100x the script's flow: at the original 0.01 m/s the T-independent k0*C^2 rate cools the adiabatic bed to ~1 K at
    # z = 1, where T = Tf*(1+theta) loses 3 digits to cancellation (an ill-conditioned scenario, not a solver property)
    sweep["VoFlRa"] = np.full(N, 100*mi["feed"]["volumetric-flowrate"])
    out = rmtExeBatch(mi, sweep, profile=True)
    assert (out["status"] == 0).all() and out["dataYs"].shape == (N, 5, 101)
    idx = [0, 150, 299]
    ens = O.Ensemble(mi, {k: v[idx] for k, v in sweep.items()})
    _, ref, _ = O.solve_n1(ens, method="BDF", rtol=1e-11, atol=1e-14)
    assert relerr(out["dataYs"][idx], ref) < RTOL_SOLVE
    mi = cases.readme_case("N1", "iso-thermal")
    rows = cases.ensemble_rows(200)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"]}
    out = rmtExeBatch(mi, sweep, profile=True)
    assert (out["status"] == 0).all() and out["dataYs"].shape == (200, 7, 101)
    idx = [3, 199]
    ens = O.Ensemble(mi, {k: v[idx] for k, v in sweep.items()})
    _, ref, _ = O.solve_n1(ens, method="BDF", rtol=1e-11, atol=1e-14)
    assert relerr(out["dataYs"][idx], ref) < RTOL_SOLVE


def test_parameter_estimation_forward_sensitivities(torch_cuda):
    """SURVEY.md section 8 f2 / BASELINE config 5 without finite differences: d(outlet)/d(theta) [x_i, P, T] for the six kinetic constants
    from K2s (exact sensitivity equations), against central differences of the oracle (SciPy at converged tolerance) on 4
    reactors, and against the forward-difference path on 512."""
    import parity_helpers as H
    from pyremot_b200 import estimation
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    rows = cases.ensemble_rows(512)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    J, base = estimation.outlet_jacobian_sens(mi, sweep, theta, rtol=1e-10, atol=1e-14)
    assert J.shape == (512, 8, 6) and np.isfinite(J).all()
    Jfd, base_fd = estimation.outlet_jacobian_fd(mi, sweep, theta, rel=1e-4)
    assert relerr(base, base_fd) < 1e-8
    for k, name in enumerate(theta):
        scale = np.abs(Jfd[:, :6, k]).max(axis=1)
        # forward differences: O(rel) truncation error, plus the noise of two differenced solves (rtol 1e-11) over the step --
        # which is all there is in the K3 column (the outlet sits at that reaction's equilibrium, d/dK3 ~ 1e-12)
        floor = 5e-11*np.abs(base[:, :6]).max(axis=1)/(1e-4*cases.MEOH_DME_THETA[name])
        assert (np.abs(J[:, :6, k] - Jfd[:, :6, k]).max(axis=1) < 3e-3*scale + floor).all(), k
    idx = [17, 100, 300, 511]
    nb, rel = len(idx), 1e-5
    sw = {k: np.concatenate([v[idx]]*12) for k, v in sweep.items()}
    for k, name in enumerate(theta):
        col = np.full(12*nb, cases.MEOH_DME_THETA[name])
        col[(2*k)*nb:(2*k + 1)*nb] *= 1 + rel
        col[(2*k + 1)*nb:(2*k + 2)*nb] *= 1 - rel
        sw["VARS:" + name] = col
    ref = H.oracle_n1(sw, range(12*nb), rtol=1e-13, atol=1e-17, theta_as_vars=True).reshape(6, 2, nb, 8)
    worst = 0.0
    for k, name in enumerate(theta):
        Jref = (ref[k, 0] - ref[k, 1])/(2*rel*cases.MEOH_DME_THETA[name])            # [nb, 8]
        for b, i in enumerate(idx):
            scale = np.abs(Jref[b, :6]).max()
            floor = 2e-13*np.abs(base[i, :6]).max()/(2*rel*cases.MEOH_DME_THETA[name])      # noise of the differenced oracle solves
            dev = np.abs(J[i, :6, k] - Jref[b, :6]).max()
            worst = max(worst, dev/(2e-5*scale + floor))
            assert dev < 2e-5*scale + floor, (i, k, dev, scale, floor)
            assert abs(J[i, 7, k] - Jref[b, 7]) < 2e-5*abs(Jref[b, 7]) + 2e-13*base[i, 7]/(2*rel*cases.MEOH_DME_THETA[name])
    print(f"\nforward sensitivities vs oracle central differences: worst deviation {worst:.2f} of the allowed band (2e-5 relative)")


def test_rmtExe_m7_matches_converged_reference(torch_cuda, golden):
    """rmtExe with model "M7" (dimensional steady-state balances, `rmtCore.py:300-309` -> `pbReactor.py:1170`) on the
    GPU: result layout of the reference (dataYs = [x_i ; T] over 30 points, XYList, dataList), profile within 1e-6 of the
    reference's converged run; an 8-row sweep through rmtExeBatch against reference-made converged profiles."""
    from pyremot_b200 import rmtExe, rmtExeBatch
    g = golden("m7")
    res = rmtExe(cases.m7_case())["resModel"]
    assert set(res) == {"dataYs", "XYList", "dataList"} and res["dataYs"].shape == (7, 30)
    assert [d["leg"] for d in res["dataList"]] == list(g["default_legs"]) and len(res["XYList"]) == 7
    assert np.allclose(res["XYList"][0][0], g["default_X"]) and np.array_equal(res["dataList"][3]["y"], res["dataYs"][3])
    assert relerr(res["dataYs"], g["conv_dataYs"]) < RTOL_SOLVE
    assert relerr(res["dataYs"], g["default_dataYs"]) < 2e-3
    rows = cases.ensemble_rows(8)
    out = rmtExeBatch(cases.m7_case(), {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}, profile=True)
    assert (out["status"] == 0).all() and out["dataYs"].shape == (8, 8, 30)
    got = np.concatenate([out["dataYs"][:, :6], out["dataYs"][:, 7:8]], axis=1)      # [x_i ; T] (the batch entry keeps the P row)
    err = relerr(got, g["sweep_conv_dataYs"])
    print(f"\nM7: 8 swept profiles vs the reference's converged runs: worst rel. error {err:.2e}")
    assert err < RTOL_SOLVE


def test_rmtExe_m8_matches_converged_reference(torch_cuda, golden):
    """rmtExe with model "M8" (velocity as an unknown, the reference's built-in kinetics; `rmtCore.py:311-320` ->
    `pbReactor.py:1580`) on the GPU: result layout of the reference (dataYs = [x_i ; T] over 25 points; XYList / dataList with the
    n+3 series x_i, T, P, v), all series within 1e-6 of the reference's converged run; an 8-row sweep through rmtExeBatch."""
    from pyremot_b200 import rmtExe, rmtExeBatch
    g = golden("m8")
    res = rmtExe(cases.m8_case())["resModel"]
    assert set(res) == {"dataYs", "XYList", "dataList"} and res["dataYs"].shape == (7, 25)
    assert [d["leg"] for d in res["dataList"]] == list(g["default_legs"]) and len(res["XYList"]) == 9
    series = np.array([xy[1] for xy in res["XYList"]])
    assert np.allclose(res["XYList"][0][0], g["default_X"]) and np.array_equal(series[:7], res["dataYs"])
    assert relerr(series, g["conv_series"]) < RTOL_SOLVE
    assert relerr(res["dataYs"], g["default_dataYs"]) < 2e-3
    mi = cases.m8_case()
    mi["reaction-rates"] = {"VARS": {}, "RATES": {"r1": lambda x: 0.0, "r2": lambda x: 0.0, "r3": lambda x: 0.0}}
    assert np.array_equal(rmtExe(mi)["resModel"]["dataYs"], res["dataYs"])       # like the reference, M8 ignores the user's rates
    rows = cases.ensemble_rows(8)
    out = rmtExeBatch(cases.m8_case(), {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}, profile=True)
    assert (out["status"] == 0).all() and out["dataYs"].shape == (8, 9, 25)
    got = out["dataYs"][:, [0, 1, 2, 3, 4, 5, 7, 6, 8]]                         # batch rows are [x_i ; P ; T ; v]
    err = relerr(got, g["sweep_conv_series"])
    print(f"\nM8: 8 swept series [x_i, T, P, v] vs the reference's converged runs: worst rel. error {err:.2e}")
    assert err < RTOL_SOLVE
