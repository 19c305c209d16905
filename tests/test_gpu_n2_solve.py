"""GPU parity of the N2 path (K0 -> K4 -> K5b through the C-ABI) against the reference's rmtExe("N2") outputs
(tests/golden/n2_solve.npz: SciPy BDF tightened to rtol 1e-9 / atol 1e-12) and against the oracle at the
BASELINE config-3 grid (zNo = 100).  Tolerance 1e-6 relative on mole fractions and T over the whole
axial x temporal profile (north_star)."""
import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases

pytestmark = pytest.mark.gpu
RTOL_SOLVE = 1e-6


def relerr(a, b, floor=0.0):
    return float(np.max(np.abs(np.asarray(a, float) - b)/(np.abs(b) + floor)))


@pytest.mark.parametrize("tag,mk,zNo", [
    ("readme_z20", lambda: cases.readme_case("N2"), 20),
    ("dme_z20", cases.dme_n2_case, 20),                       # PyREMOT/tests/test_rmt_N1_DME.py
    ("ch4_z10", cases.ch4_n2_case, 10),                       # PyREMOT/tests/test_rmt_N2_CH4.py (adiabatic, SpCoi rate)
    ("readme_iso_z10", lambda: cases.readme_case("N2", "iso-thermal"), 10),
])
def test_rmtExe_n2_matches_converged_reference(golden, tag, mk, zNo):
    import pyremot_b200
    from pyremot_b200 import rmtExe
    g = golden("n2_solve")
    saved = dict(pyremot_b200.solverSetting["N2"])
    try:
        pyremot_b200.solverSetting["N2"]["zNo"] = zNo         # same mechanism as PyREMOT.solvers.solSetting
        res = rmtExe(mk())["resModel"]
    finally:
        pyremot_b200.solverSetting["N2"].update(saved)
    assert set(res) == {"computation-time", "dataPack"} and len(res["dataPack"]) == 5
    for k, d in enumerate(res["dataPack"]):
        assert set(d) == {"modelId", "processType", "successStatus", "dataShape", "labelList", "indexList", "dataTime",
                          "dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2", "dataYs"}
        assert d["dataShape"] == () and d["labelList"][-1] == "Temperature"
        assert abs(d["dataTime"] - g[f"{tag}_conv_dataTime"][k]) < 1e-12
        for key in ("dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2", "dataYs"):
            ref = g[f"{tag}_conv_{key}"][k] if key != "dataXs" else g[f"{tag}_conv_dataXs"]
            assert np.shape(d[key]) == ref.shape, (key, np.shape(d[key]), ref.shape)
            # theta = (T-Tf)/Tf starts at 0: compare it on the scale of 1 (i.e. as T/Tf)
            floor = 1.0 if key == "dataYTemp1" else (1e-9 if key == "dataXs" else 0.0)
            assert relerr(d[key], ref, floor=floor) < RTOL_SOLVE, (key, k)
        # distance to the reference's own default-tolerance output (it integrates at rtol 1e-3)
        assert relerr(d["dataYs"], g[f"{tag}_default_dataYs"][k]) < 2e-2


def test_n2_baseline_grid_vs_oracle():
    """BASELINE config 3 shape (zNo = 100, tNo = 5, period 0.5 s, start-up from the feed-filled bed): 64 swept
    instances on the GPU; two of them against the oracle (SciPy BDF on the NumPy RHS, rtol 1e-9)."""
    import torch
    from pyremot_b200.api import get_model
    N, zNo, tNo = 64, 100, 5
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    tab = m.instance_table(mi, sweep)
    r = m.solve_n2_host(tab, zNo, tNo, 0.5)
    assert (r["status"] == 0).all()
    x = r["dataYs"][:, :, :6, :]
    assert np.max(np.abs(x.sum(axis=2) - 1)) < 1e-12
    idx = [0, 37]
    ens = O.Ensemble(mi, {k: v[idx] for k, v in sweep.items()})
    _, ref, _ = O.solve_n2(ens, zNo=zNo, tNo=tNo, method="BDF", rtol=1e-9, atol=1e-13, sparse=True)
    assert relerr(r["dataYs"][idx], ref) < RTOL_SOLVE
    # device-pointer path == host-buffer path; sub-batch reproduces the same rows; ragged grid (zNo not a multiple of 32)
    par = m.prolog(torch.from_numpy(tab).cuda())
    rd = m.solve_n2(par, zNo, tNo, 0.5)
    assert np.array_equal(m.unscale_n2(par, rd["snaps_y"]).cpu().numpy(), r["dataYs"])
    r2 = m.solve_n2_host(tab[10:13], zNo, tNo, 0.5)
    assert np.array_equal(r2["dataYs"], r["dataYs"][10:13])


def test_n2_more_instances_than_resident_warps():
    """N larger than the persistent grid (148 SMs x 8 warps): every warp loops over several instances."""
    from pyremot_b200.api import get_model
    N, zNo = 3000, 12
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    tab = m.instance_table(mi, sweep)
    r = m.solve_n2_host(tab, zNo, 2, 0.2)
    assert (r["status"] == 0).all()
    sub = m.solve_n2_host(tab[2500:2510], zNo, 2, 0.2)
    assert np.array_equal(sub["dataYs"], r["dataYs"][2500:2510])
    ens = O.Ensemble(mi, {k: v[[2999]] for k, v in sweep.items()})
    ens.period = 0.2
    _, ref, _ = O.solve_n2(ens, zNo=zNo, tNo=2, method="BDF", rtol=1e-9, atol=1e-13, sparse=True)
    assert relerr(r["dataYs"][[2999]], ref) < RTOL_SOLVE


def test_rmtExeBatch_n2_and_failure_status():
    import pyremot_b200
    from pyremot_b200 import rmtExeBatch
    mi = cases.readme_case("N2")
    rows = cases.ensemble_rows(8)
    sweep = {"T": rows["T"].copy(), "P": rows["P"].copy(), "conc": rows["conc"].copy(), "Tm": rows["T"].copy()}
    saved = dict(pyremot_b200.solverSetting["N2"])
    try:
        pyremot_b200.solverSetting["N2"].update({"zNo": 16, "tNo": 2})
        good = rmtExeBatch(mi, sweep)
        sweep["T"][3] = float("nan")
        out = rmtExeBatch(mi, sweep)
    finally:
        pyremot_b200.solverSetting["N2"].update(saved)
    assert good["dataYs"].shape == (8, 2, 7, 16) and (good["status"] == 0).all()
    assert out["status"][3] != 0 and (np.delete(out["status"], 3) == 0).all()
    assert np.array_equal(np.delete(out["dataYs"], 3, axis=0), np.delete(good["dataYs"], 3, axis=0))


def test_n2_both_kernels_agree_and_large_grid():
    """K4b (block per reactor, zNo <= 256) and K4 (warp per reactor, any zNo) integrate the same equations: they must agree
    to integration tolerance; zNo = 400 (BASELINE config 4 grid) runs through K4d and K4c and is checked against the oracle."""
    import os
    from pyremot_b200.api import get_model
    mi = cases.readme_case("N2")
    m = get_model(mi)
    rows = cases.ensemble_rows(6)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    tab = m.instance_table(mi, sweep)
    blk = m.solve_n2_host(tab, 64, 2, 0.2)
    os.environ["PBR_N2_KERNEL"] = "warp"
    try:
        wrp = m.solve_n2_host(tab, 64, 2, 0.2)
    finally:
        del os.environ["PBR_N2_KERNEL"]
    assert (blk["status"] == 0).all() and (wrp["status"] == 0).all()
    assert relerr(blk["dataYs"], wrp["dataYs"]) < 1e-7
    big = m.solve_n2_host(tab[:2], 400, 1, 0.05)             # K4d: all on chip, scan march
    assert (big["status"] == 0).all()
    os.environ["PBR_K4C_ONCHIP"] = "0"                       # K4c: history in a global slab (what a model whose state
    try:                                                     # does not fit on chip gets at this grid size)
        slab = m.solve_n2_host(tab[:2], 400, 1, 0.05)
    finally:
        del os.environ["PBR_K4C_ONCHIP"]
    assert (slab["status"] == 0).all() and relerr(slab["dataYs"], big["dataYs"]) < 1e-7
    ens = O.Ensemble(mi, {k: v[:1] for k, v in sweep.items()})
    ens.period = 0.05
    _, ref, _ = O.solve_n2(ens, zNo=400, tNo=1, method="BDF", rtol=1e-9, atol=1e-13, sparse=True)
    assert relerr(big["dataYs"][:1], ref) < RTOL_SOLVE


def test_n2_scan_march_matches_serial_march_and_oracle():
    """Grids above 56 nodes use the log-depth scan of the per-node affine maps (pbr_k4bs_solve_n2) instead of the serial
    node march (pbr_k4b_solve_n2): both solve the same block-bidiagonal Newton systems, so the snapshots must agree to
    integration tolerance, and with the oracle to 1e-6.  zNo = 160 is ragged (5 warps of nodes)."""
    import os
    from pyremot_b200.api import get_model
    mi = cases.readme_case("N2")
    m = get_model(mi)
    rows = cases.ensemble_rows(5)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    tab = m.instance_table(mi, sweep)
    scan = m.solve_n2_host(tab, 160, 2, 0.1)
    os.environ["PBR_K4B_SCAN_FROM"] = "512"
    try:
        serial = m.solve_n2_host(tab, 160, 2, 0.1)
    finally:
        del os.environ["PBR_K4B_SCAN_FROM"]
    assert (scan["status"] == 0).all() and (serial["status"] == 0).all()
    assert relerr(scan["dataYs"], serial["dataYs"]) < 1e-7
    ens = O.Ensemble(mi, {k: v[:1] for k, v in sweep.items()})
    ens.period = 0.1
    _, ref, _ = O.solve_n2(ens, zNo=160, tNo=2, method="BDF", rtol=1e-9, atol=1e-13, sparse=True)
    assert relerr(scan["dataYs"][:1], ref) < RTOL_SOLVE


def test_config3_full_size_hardest_instances_vs_oracle():
    """BASELINE configs[2] at full size: 4,096 reactors x 100 nodes, tNo = 5, period 0.5 s.  Every instance integrates; all five
    snapshots (full axial profiles) of the 8 reactors that needed the most RHS evaluations and of 8 seeded ones are compared with
    the oracle (SciPy BDF rtol 1e-9 on the NumPy RHS, `pbHomoReactor.py:3589-3685` interval loop)."""
    import torch
    import parity_helpers as H
    from pyremot_b200.api import get_model
    N, zNo, tNo, period = 4096, 100, 5, 0.5
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    par = m.prolog(torch.from_numpy(m.instance_table(mi, sweep)).cuda())
    r = m.solve_n2(par, zNo, tNo, period)
    assert int((r["status"] != 0).sum()) == 0
    nfev = r["stats"][2].cpu().numpy()
    idx = sorted(set(np.argsort(nfev)[-8:].tolist() + np.random.default_rng(11).choice(N, 8, replace=False).tolist()))
    sel = torch.tensor(idx, device="cuda")
    data = m.unscale_n2(par[:, sel].contiguous(), r["snaps_y"][sel].contiguous()).cpu().numpy()
    ref = H.oracle_n2(sweep, idx, zNo, tNo, period)
    err = H.relerr_rows(data, ref)
    print(f"\nconfig 3 (N = 4096): {len(idx)} reactors x 5 snapshots x 100 nodes vs oracle: worst rel. error {err.max():.2e} "
          f"(max nfev {nfev.max()}, mean {nfev.mean():.0f})")
    assert err.max() < RTOL_SOLVE/3, err


def test_config4_grid_full_period_vs_oracle():
    """BASELINE configs[3] grid (400 nodes) over the FULL transient (tNo = 5, period 0.5 s) through K4d: 4 reactors, all snapshots
    against the oracle."""
    import parity_helpers as H
    from pyremot_b200.api import get_model
    N, zNo, tNo, period = 4, 400, 5, 0.5
    rows = cases.ensemble_rows(N, offset=4242)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    r = m.solve_n2_host(m.instance_table(mi, sweep), zNo, tNo, period)
    assert (r["status"] == 0).all()
    ref = H.oracle_n2(sweep, range(N), zNo, tNo, period)
    err = H.relerr_rows(r["dataYs"], ref)
    print(f"\nconfig-4 grid (zNo = 400, full period): worst rel. error {err.max():.2e}, steps {r['stats'][0].tolist()}")
    assert err.max() < RTOL_SOLVE/3, err


def test_warp_kernel_on_a_grid_that_needs_it_vs_oracle():
    """Grids above 512 nodes run through K4 (one warp per reactor, scratch in global memory): zNo = 640 against the oracle."""
    import parity_helpers as H
    from pyremot_b200.api import get_model
    N, zNo, tNo, period = 3, 640, 2, 0.06
    rows = cases.ensemble_rows(N, offset=99)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    r = m.solve_n2_host(m.instance_table(mi, sweep), zNo, tNo, period)
    assert (r["status"] == 0).all()
    ref = H.oracle_n2(sweep, range(N), zNo, tNo, period)
    err = H.relerr_rows(r["dataYs"], ref)
    print(f"\nwarp kernel (zNo = 640): worst rel. error {err.max():.2e}")
    assert err.max() < RTOL_SOLVE/3, err


@pytest.mark.parametrize("zNo", [100, 400, 640])
def test_resume_from_saved_snapshot(zNo):
    """intervals 0..1 then a resume at interval 2 from the saved snapshot (the reference's restart,
    IV = dataYs[:, -1], `pbHomoReactor.py:3685`) reproduce the uninterrupted run -- to 1e-9 when the integrator restarts at every
    interval as the reference does, within the parity bar otherwise.  All three kernel families (K4b, K4d, K4 warp)."""
    import torch
    from pyremot_b200.api import get_model
    N, tNo, period, k = 3, 5, 0.1, 2
    rows = cases.ensemble_rows(N, offset=7)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    tab = m.instance_table(mi, sweep)
    par = m.prolog(torch.from_numpy(tab).cuda())
    for restart, tol in ((True, 1e-9), (False, RTOL_SOLVE)):
        full = m.solve_n2(par, zNo, tNo, period, restart_every_interval=restart)
        assert int((full["status"] != 0).sum()) == 0
        tail = m.solve_n2(par, zNo, tNo, period, iv=full["snaps_y"][:, k - 1].contiguous(), k_first=k, restart_every_interval=restart)
        assert int((tail["status"] != 0).sum()) == 0 and tuple(tail["snaps_y"].shape) == (N, tNo - k, m.nv2, zNo)
        a, b = tail["snaps_y"].cpu().numpy(), full["snaps_y"][:, k:].cpu().numpy()
        assert relerr(a, b, floor=1e-9) < tol, (zNo, restart, relerr(a, b, floor=1e-9))
    # host-buffer entry point, and a state that cannot come from this model
    fh = m.solve_n2_host(tab, zNo, tNo, period)
    th = m.solve_n2_host(tab, zNo, tNo, period, iv=fh["snaps_y"][:, k - 1], k_first=k)
    assert (th["status"] == 0).all() and relerr(th["dataYs"], fh["dataYs"][:, k:]) < RTOL_SOLVE
    if zNo <= 512:                                       # the block kernels integrate in extent coordinates and refuse it
        bad = fh["snaps_y"][:, k - 1].copy()
        bad[1, 0] *= 1.001
        tb = m.solve_n2_host(tab, zNo, tNo, period, iv=bad, k_first=k)
        assert tb["status"][1] == 5 and np.isnan(tb["dataYs"][1]).all() and (tb["status"][[0, 2]] == 0).all()


def test_saved_shards_are_resume_points(tmp_path):
    """SURVEY.md section 8 f4 end to end: a dynamic ensemble run is written with `io.save_shard` (two ranks' worth of rows), read back with
    `io.load_shards`, and continued from its LAST snapshot over a second period -- the same states as continuing in memory."""
    from pyremot_b200 import io
    from pyremot_b200.api import get_model
    N, zNo, tNo, period = 6, 100, 2, 0.04
    rows = cases.ensemble_rows(N, offset=21)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    tab = m.instance_table(mi, sweep)
    first = m.solve_n2_host(tab, zNo, tNo, period)
    assert (first["status"] == 0).all()
    for rank, (lo, hi) in enumerate(((0, 4), (4, 6))):
        io.save_shard(str(tmp_path), {k: (v[:, lo:hi] if k == "stats" else v[lo:hi]) for k, v in first.items()}, lo, hi, rank=rank,
                      meta={"zNo": zNo, "tNo": tNo, "period": period})
    saved, meta = io.load_shards(str(tmp_path))
    assert meta["zNo"] == zNo and np.array_equal(saved["snaps_y"], first["snaps_y"])
    # the run continues on a grid of 2*tNo intervals over 2*period: intervals tNo .. 2*tNo-1 start from the saved last snapshot
    cont = m.solve_n2_host(tab, zNo, 2*tNo, 2*period, iv=saved["snaps_y"][:, -1], k_first=tNo)
    whole = m.solve_n2_host(tab, zNo, 2*tNo, 2*period)
    assert (cont["status"] == 0).all() and cont["dataYs"].shape == (N, tNo, 7, zNo)
    assert np.allclose(whole["dataYs"][:, :tNo], first["dataYs"], rtol=1e-12, atol=0)        # same first half
    assert relerr(cont["dataYs"], whole["dataYs"][:, tNo:]) < RTOL_SOLVE
