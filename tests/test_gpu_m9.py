"""GPU parity of model "M9" (`PyREMOT/docs/rmtCore.py:322-331` -> `pbReactor.py` runM5 `:1997`,
modelEquationM5 `:2296-2666`), the dynamic dimensional model whose pressure AND velocity are marched along the bed inside every
evaluation.  It runs through the N2 entry points of the C-ABI: the block-per-reactor kernels up to 512 nodes (block-wide fixed-point
sweeps of the march, pbr_k4b_eval) and the warp-per-reactor kernel above and for the RHS alone (per chunk of 32 nodes,
pbr_n2_sweep_eval) -- grids of more than one chunk and ragged chunks are the cases that only the GPU exercises for the latter (the
host emulation of the warp kernel has one lane)."""
import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases

pytestmark = pytest.mark.gpu


def _scaled(par, Y):
    """dimensional [C_i ; T] -> the solver's y = [C_i/Cmax ; (T - T0)/T0]"""
    Cmax, T0 = par[2][:, None, None], par[0][:, None]
    y = Y.copy()
    y[:, :-1] /= Cmax
    y[:, -1] = (Y[:, -1] - T0)/T0
    return y


def test_m9_rhs_vs_reference_golden_and_oracle(golden):
    import torch
    from pyremot_b200.api import get_model
    g = golden("m9")
    mi = cases.m9_case()
    m = get_model(mi)
    assert m.variant == "M9"
    for zNo in (12, 40):                                        # the reference's own RHS on random states (one / two ragged chunks)
        Y, F = g[f"rhs_y_z{zNo}"], g[f"rhs_f_z{zNo}"]
        par = m.prolog(torch.from_numpy(m.instance_table(mi, N=len(Y))).cuda())
        ph = par.cpu().numpy()
        f = m.rhs_n2(par, torch.from_numpy(_scaled(ph, Y)).cuda()).cpu().numpy()
        f[:, :-1] *= ph[2][:, None, None]
        f[:, -1] *= ph[0][:, None]
        assert np.max(np.abs(f - F)/(np.abs(F) + 1e-12*np.abs(F).max())) < 1e-10, zNo
    # the default grid of solverSetting['S2'] (100 nodes: three full chunks and a ragged one) and an odd one, swept instances,
    # against the oracle (itself pinned to the reference's RHS at the two grids above)
    rng = np.random.default_rng(99)
    for zNo in (100, 33, 257):
        N = 96
        rows = cases.ensemble_rows(N)
        sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - 5.0}
        ens = O.Ensemble(mi, sweep)
        Y = ens.iv_m9(zNo)
        Y[:, :-1] *= np.exp(rng.normal(0, 0.05, Y[:, :-1].shape))
        Y[:, -1] += np.cumsum(rng.uniform(-0.5, 1.0, (N, zNo)), axis=1)
        F = ens.rhs_m9(Y)
        par = m.prolog(torch.from_numpy(m.instance_table(mi, dict(sweep))).cuda())
        ph = par.cpu().numpy()
        f = m.rhs_n2(par, torch.from_numpy(_scaled(ph, Y)).cuda()).cpu().numpy()
        f[:, :-1] *= ph[2][:, None, None]
        f[:, -1] *= ph[0][:, None]
        # the upwind differences (C_z - C_{z-1})/dz cancel: the absolute rounding error is ~1e-16 of the LARGEST term, so entries
        # that are small against their own row's scale carry a larger relative one -- hence the floor at 1e-6 of that scale
        assert np.max(np.abs(f - F)/(np.abs(F) + 1e-6*np.abs(F).max(axis=(1, 2), keepdims=True))) < 1e-9, zNo


def test_rmtExe_m9_matches_converged_reference(golden):
    import pyremot_b200
    from pyremot_b200 import rmtExe
    g = golden("m9")
    saved = dict(pyremot_b200.solverSetting["S2"])
    try:
        for tag, mi, zNo in (("z12", cases.m9_case(period=0.2), 12), ("z20", cases.m9_case(period=0.1, T=540.0, Tm=530.0), 20)):
            pyremot_b200.solverSetting["S2"].update({"zNo": zNo, "tNo": 2})     # same mechanism as PyREMOT.solvers.solSetting
            res = rmtExe(mi)["resModel"]
            assert {"XYList", "dataList"} <= set(res) and len(res["XYList"]) == 2
            for k in range(2):
                assert res["dataList"][k]["leg"] == str(g[f"{tag}_legs"][k])
                assert np.max(np.abs(res["XYList"][k][0] - g[f"{tag}_X"])) < 1e-12
                assert np.max(np.abs(res["XYList"][k][1] - g[f"{tag}_conv_T"][k]))/500 < 1e-9, (tag, k)
                assert np.max(np.abs(res["dataList"][k]["y"] - g[f"{tag}_default_T"][k]))/500 < 1e-6, (tag, k)
    finally:
        pyremot_b200.solverSetting["S2"].update(saved)


def test_m9_default_grid_ensemble_vs_oracle():
    """solverSetting['S2'] as the reference ships it (zNo 100, tNo 10), 2 s of start-up with the wall 10 K below the feed, 64 swept
    instances through rmtExeBatch; two of them against the oracle's converged interval loop (SciPy BDF on the NumPy RHS)."""
    import parity_helpers as PH
    from pyremot_b200 import rmtExeBatch
    N, period = 64, 2.0
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - 10.0}
    mi = cases.m9_case(period=period)
    r = rmtExeBatch(mi, dict(sweep))
    assert (r["status"] == 0).all()
    data = r["dataYs"]                                          # [N, tNo, n+1, zNo] = [x_i ; T]
    assert data.shape == (N, 10, 7, 100)
    idx = [int(np.argmax(rows["T"])), int(np.argmin(rows["P"]))]
    ref = PH.oracle_m9(sweep, idx, 100, 10, period)             # dimensional [C_i ; T]
    x_ref = ref[:, :, :-1]/ref[:, :, :-1].sum(axis=2, keepdims=True)
    for j, i in enumerate(idx):
        assert np.max(np.abs(data[i, :, -1] - ref[j, :, -1]))/500 < 1e-6, i
        assert np.max(np.abs(data[i, :, :-1] - x_ref[j])/(x_ref[j] + 1e-6)) < 1e-5, i
    # the start-up front really moves through the bed in these 2 s (a test of a frozen profile would prove little)
    assert np.abs(data[idx[0], -1, -1] - data[idx[0], 0, -1]).max() > 1.0


def test_m9_block_and_warp_kernels_agree():
    """The solver paths -- block per reactor with the serial march, block per reactor with the scan march, warp per reactor --
    integrate the same equations with the same controller: same step counts, snapshots within integration tolerance of a run at
    rtol 1e-11, at 33 / 64 / 100 nodes (ragged, two and four warps).  (Finer grids are not used here: the reference's M9 equations
    themselves turn ill-conditioned there -- at 160 nodes two converged runs of ANY of the kernels, and of the oracle, differ by 1e-6
    at the outlet node, at 300 nodes the integration stalls -- which is a property of the model, see DESIGN.md.)"""
    import os
    import torch
    from pyremot_b200.api import get_model
    mi = cases.m9_case(period=0.4)
    m = get_model(mi)
    N = 8
    rows = cases.ensemble_rows(N)
    par = m.prolog(torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - 10.0})).cuda())

    def run(zNo, kern, rtol, atol):
        env = {"warp": ("PBR_N2_KERNEL", "warp"), "serial": ("PBR_K4B_SCAN_FROM", "100000"), "scan": ("PBR_K4B_SCAN_FROM", "0")}[kern]
        os.environ[env[0]] = env[1]
        try:
            r = m.solve_n2(par, zNo, 2, 0.4, rtol=rtol, atol=atol)
        finally:
            del os.environ[env[0]]
        assert (r["status"] == 0).all(), (zNo, kern)
        return r["snaps_y"].cpu().numpy(), r["stats"].cpu().numpy()
    for zNo in (33, 64, 100):
        truth, _ = run(zNo, "warp", 1e-11, 1e-15)
        steps = None
        for kern in ("warp", "serial", "scan"):
            y, st = run(zNo, kern, 1e-9, 1e-13)
            assert np.max(np.abs(y - truth)) < 5e-8, (zNo, kern, np.max(np.abs(y - truth)))
            steps = st[0] if steps is None else steps
            assert np.array_equal(st[0], steps), (zNo, kern)
