"""The device code (csrc/pbr_device.cuh), compiled as plain C++ by tests/hostsim, against reference-generated
golden vectors — numerics of the kernels' per-thread / per-warp functions checked on the CPU-only container.
(The product never uses this build; the same checks run on the GPU through the C-ABI in tests/test_gpu_*.py.)"""
import numpy as np
import pytest

from hostsim import HostSim
from oracle import pbr_oracle as O
from pyremot_b200 import cases, engine


def relerr(a, b, floor=0.0):
    return float(np.max(np.abs(np.asarray(a, float) - b)/(np.abs(b) + floor)))


@pytest.fixture(scope="module")
def readme():
    mi = cases.readme_case("N1")
    m = engine.Model(mi, device=None)
    return mi, m, HostSim(m)


def test_prolog_and_rhs_n1_vs_reference(readme, golden):
    mi, m, hs = readme
    g = golden("n1_rhs")
    rows = cases.ensemble_rows(24)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    par = hs.prolog(m.instance_table(mi, sweep))
    ens = O.Ensemble(mi, sweep)
    for row, ref in ((0, ens.T0), (2, ens.Cmax), (4, ens.vs0), (9, ens.rho0), (10, ens.Cpmix0), (11, ens.G), (12, ens.GH), (15, ens.mu)):
        assert relerr(par[row], ref) < 1e-14
    for k in range(2):
        y = np.ascontiguousarray(g["y_sweep"][:, k].T)
        f, J = hs.jac_n1(par, y)
        assert relerr(f.T, g["f_sweep"][:, k], floor=1e-300) < 1e-12
        assert np.array_equal(hs.rhs_n1(par, y), f)
        if k == 1:
            for i in (0, 11):
                e = ens.sub(i)
                Jfd = O.fd_jacobian(lambda yy: e.rhs_n1(yy[None])[0], y[:, i], rel=1e-6)
                assert np.max(np.abs(J[:, :, i] - Jfd)/np.max(np.abs(Jfd), axis=0, keepdims=True)) < 1e-5
    par1 = hs.prolog(m.instance_table(mi, N=600))
    assert relerr(hs.rhs_n1(par1, np.ascontiguousarray(g["y_noniso"].T)).T, g["f_noniso"], floor=1e-300) < 1e-12


@pytest.mark.parametrize("tag,mk", [("readme_z12", lambda: cases.readme_case("N2")), ("dme_z20", cases.dme_n2_case),
                                    ("ch4_z10", cases.ch4_n2_case), ("readme_iso_z8", lambda: cases.readme_case("N2", "iso-thermal"))])
def test_rhs_n2_vs_reference(golden, tag, mk):
    g = golden("n2_rhs")
    mi = mk()
    m = engine.Model(mi, device=None)
    hs = HostSim(m)
    Y, F = g[f"y_{tag}"], g[f"f_{tag}"]
    par = hs.prolog(m.instance_table(mi, N=len(Y)))
    assert relerr(hs.rhs_n2(par, Y), F, floor=1e-9) < 1e-11


def test_bdf_n1_vs_converged_reference(readme, golden):
    mi, m, hs = readme
    gs = golden("n1_sweep")
    sweep = {"T": gs["T"], "P": gs["P"], "conc": gs["conc"], "Tm": gs["T"]}
    par = hs.prolog(m.instance_table(mi, sweep))
    prof, status, stats = hs.solve_n1(par, 1e-9, 1e-13)
    assert (status == 0).all()
    ys = O.Ensemble(mi, sweep).unscale_n1(np.transpose(prof, (2, 1, 0)))
    assert relerr(ys, gs["dataYs"]) < 1e-6
    assert stats[4].max() <= 2 and stats[0].max() < 900        # Newton failures are rare; ~600 steps per reactor


def test_bdf_n2_vs_converged_reference(golden):
    g = golden("n2_solve")
    for tag, mk, zNo in (("dme_z20", cases.dme_n2_case, 20), ("ch4_z10", cases.ch4_n2_case, 10)):
        mi = mk()
        m = engine.Model(mi, device=None)
        hs = HostSim(m)
        ens = O.Ensemble(mi)
        snaps, status, stats = hs.solve_n2(hs.prolog(m.instance_table(mi)), zNo, 5, float(ens.period), 1e-9, 1e-13)
        assert status[0] == 0
        data = np.stack([ens.unscale_n2(snaps[:, k]) for k in range(5)], axis=1)[0]
        assert relerr(data, g[f"{tag}_conv_dataYs"]) < 1e-6


def test_lu_and_history_rescale_primitives(readme):
    import ctypes as C
    from scipy.integrate._ivp.bdf import change_D
    mi, m, hs = readme
    rng = np.random.default_rng(0)
    for t in range(6):
        A, b = rng.normal(size=(8, 8)), rng.normal(size=8)
        if t == 1:
            A[0, 0] = 0.0                                   # forces a row swap in the first column
        if t == 2:
            A[:, 0] *= 1e-9
            A[3, 0] = 5.0
        x = np.linalg.solve(A, b)
        A2, b2 = A.copy(), b.copy()
        hs.L.hs_lu_test(0, A2.ctypes.data_as(C.c_void_p), b2.ctypes.data_as(C.c_void_p))
        assert np.max(np.abs(b2 - x))/np.max(np.abs(x)) < 1e-12
    for order in range(1, 6):
        for factor in (0.5, 0.2, 1.7, 10.0, 0.913):
            D = rng.normal(size=(7, 8))
            D1, D2 = D.copy(), D.copy()
            change_D(D1, order, factor)
            hs.L.hs_change_D(D2.ctypes.data_as(C.c_void_p), C.c_int(order), C.c_double(factor))
            assert np.max(np.abs(D1 - D2)) < 1e-12*np.max(np.abs(D1))


def _extent_map(m, nv, nx):
    """M = d[c, (p), theta]/dx for the extent coordinates x = [x_1..x_R, (p), theta]."""
    R, n = m.n_rxn, m.n
    M = np.zeros((nv, nx))
    M[:n, :R] = np.asarray(m.S).reshape(R, n).T
    for k in range(nv - n):
        M[n + k, R + k] = 1.0
    return M


def test_n1_extent_coordinates_match_full_system(readme):
    """K2 integrates x = [extents, p, theta] (c = c0 + S^T x): its RHS / Jacobian must be the full system's, f = M f_x and
    M J_x = J M, and the integration must give the same profiles with the same step counts as the golden sweep test."""
    mi, m, hs = readme
    nx = hs.L.hs_nx()
    assert nx == m.n_rxn + 2 < hs.nv1                      # reduced: 5 instead of 8 variables
    rows = cases.ensemble_rows(16)
    par = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    N, R = par.shape[1], m.n_rxn
    rng = np.random.default_rng(0)
    x = np.zeros((nx, N))
    x[:R] = rng.uniform(0, 0.02, (R, N)); x[R] = rng.uniform(0.8, 1, N); x[R + 1] = rng.uniform(0, 0.1, N)
    y, fx, Jx = hs.n1x(par, x)
    f, J = hs.jac_n1(par, y)
    M = _extent_map(m, hs.nv1, nx)
    assert relerr(M @ fx, f, floor=1e-300) < 1e-12
    for i in range(N):
        assert np.max(np.abs(M @ Jx[:, :, i] - J[:, :, i] @ M)) < 1e-12*np.max(np.abs(J[:, :, i]))


@pytest.mark.parametrize("mk", [lambda: cases.readme_case("N2"), cases.ch4_n2_case, lambda: cases.readme_case("N2", "iso-thermal")])
def test_n2_extent_node_matches_full_node(mk):
    """The per-node function of K4b/K4c in extent coordinates against the full-coordinate node (pbr_n2_node)."""
    import ctypes as C
    mi = mk()
    m = engine.Model(mi, device=None)
    hs = HostSim(m)
    L = hs.L
    nx, nv, R = L.hs_nx2(), hs.nv2, m.n_rxn
    par = np.ascontiguousarray(hs.prolog(m.instance_table(mi, N=1))[:, 0])
    M = _extent_map(m, nv, nx)
    L.hs_n2x_node.restype = C.c_double
    L.hs_n2_node.restype = C.c_double
    P = HostSim._p
    rng = np.random.default_rng(1)
    for first in (0, 1):
        for _ in range(10):
            x, xb = np.zeros(nx), np.zeros(nx)
            x[:R] = rng.uniform(-2e-6, 2e-6, R); xb[:R] = rng.uniform(-2e-6, 2e-6, R)     # concentrations stay positive
            if nx > R:
                x[R], xb[R] = rng.uniform(0, 0.2, 2)
            yf, ybf, f, J, sub = np.zeros(nv), np.zeros(nv), np.zeros(nx), np.zeros((nx, nx)), np.zeros(nx)
            s1 = L.hs_n2x_node(P(par), P(x), P(xb), first, C.c_double(4.9e6), C.c_double(19.0), P(yf), P(ybf), P(f), P(J), P(sub))
            f2, J2, sub2 = np.zeros(nv), np.zeros((nv, nv)), np.zeros(nv)
            s2 = L.hs_n2_node(P(par), P(yf), P(ybf), first, C.c_double(4.9e6), C.c_double(19.0), P(f2), P(J2), P(sub2))
            assert s1 == s2
            assert relerr(M @ f, f2, floor=1e-300) < 1e-11
            assert np.max(np.abs(M @ J - J2 @ M)) < 1e-12*np.max(np.abs(J2))
            assert np.max(np.abs(M @ np.diag(sub) - np.diag(sub2) @ M)) < 1e-12*max(1.0, np.max(np.abs(sub2)))


# ---- the block-per-reactor N2 solver kernels, run whole through the SIMT emulator (tests/hostsim/simt.h)
def _simt_case(mk, extra_defines=None):
    import os
    from hostsim import SimtSim
    mi = mk()
    old = os.environ.get("PBR_EXTRA_DEFINES")
    if extra_defines:
        os.environ["PBR_EXTRA_DEFINES"] = extra_defines
    try:
        m = engine.Model(mi, device=None)
    finally:
        if extra_defines:
            if old is None:
                del os.environ["PBR_EXTRA_DEFINES"]
            else:
                os.environ["PBR_EXTRA_DEFINES"] = old
    return mi, m, HostSim(m), SimtSim(m)


def test_block_kernels_on_the_simt_emulator_vs_converged_reference(golden):
    """pbr_k4b_solve_n2 (serial march), pbr_k4bs_solve_n2 (scan march) and pbr_k4c_solve_n2 (history in a global slab), executed
    thread by thread with real barriers and shuffles, against the reference's converged rmtExe N2 snapshots (DME case, zNo = 20)
    and against the scalar host build of the warp kernel (same step counts)."""
    g = golden("n2_solve")
    mi, m, hs, sim = _simt_case(cases.dme_n2_case)
    ens = O.Ensemble(mi)
    par = hs.prolog(m.instance_table(mi))
    ref_snaps, _, ref_stats = hs.solve_n2(par, 20, 5, float(ens.period), 1e-9, 1e-13)
    for kernel in ("pbr_k4b_solve_n2", "pbr_k4bs_solve_n2", "pbr_k4c_solve_n2"):
        snaps, status, stats = sim.solve_n2(kernel, par, 20, 5, float(ens.period))
        assert status[0] == 0
        data = np.stack([ens.unscale_n2(snaps[:, k]) for k in range(5)], axis=1)[0]
        assert relerr(data, g["dme_z20_conv_dataYs"]) < 1e-6
        assert relerr(snaps, ref_snaps, floor=1e-12) < 1e-9 and abs(int(stats[0, 0]) - int(ref_stats[0, 0])) <= 2


@pytest.mark.parametrize("defines", [None, "PBR_N2_RESTART=1"])
def test_block_kernels_multi_warp_ragged_grid_and_instance_loop(defines):
    """zNo = 40: two warps, the second one ragged (8 nodes), so the marches hand over between warps, the scan chains warp totals
    and the block reductions span warps; 2 reactors on a grid of 1 block exercise the kernels' loop over instances.  All
    kernels must agree with the scalar host build of the warp kernel to integration tolerance."""
    mi, m, hs, sim = _simt_case(lambda: cases.readme_case("N2"), defines)
    rows = cases.ensemble_rows(2)
    par = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    zNo, tNo, period = 40, 2, 0.03
    ref, st0, _ = hs.solve_n2(par, zNo, tNo, period, 1e-9, 1e-13)
    assert (st0 == 0).all()
    # the reference's restart at every snapshot time (PBR_N2_RESTART=1): same body in all kernels, one is enough
    kernels = ("pbr_k4b_solve_n2",) if defines else ("pbr_k4b_solve_n2", "pbr_k4bs_solve_n2", "pbr_k4c_solve_n2", "pbr_k4d_solve_n2")
    for kernel in kernels:
        snaps, status, stats = sim.solve_n2(kernel, par, zNo, tNo, period, grid=1)
        assert (status == 0).all() and (stats[0] > 20).all()
        assert relerr(snaps, ref, floor=1e-9) < 1e-7


def test_n2_resume_from_a_saved_snapshot_on_the_simt_emulator():
    """pbr_solve_n2_resume semantics (reference restart `pbHomoReactor.py:3685`, IV = dataYs[:, -1]) on the block
    kernel, CPU-side: intervals 0..1 then a resume at interval 2 from the saved snapshot must reproduce the uninterrupted 5-interval
    run -- to rounding when the integrator restarts at every interval (flags bit 0, the reference's behaviour), to integration
    tolerance otherwise; an off-manifold state is refused with PBR_INST_BAD_IV and NaN snapshots."""
    mi, m, hs, sim = _simt_case(lambda: cases.readme_case("N2"))
    rows = cases.ensemble_rows(2)
    par = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    zNo, tNo, period, k = 40, 5, 0.05, 2
    for flags, tol in ((1, 1e-11), (0, 1e-6)):
        full, st, _ = sim.solve_n2("pbr_k4b_solve_n2", par, zNo, tNo, period, flags=flags)
        assert (st == 0).all()
        tail, st2, _ = sim.solve_n2("pbr_k4b_solve_n2", par, zNo, tNo, period, iv=full[:, k - 1], k_first=k, flags=flags)
        assert (st2 == 0).all() and tail.shape == (2, tNo - k, m.nv2, zNo)
        assert relerr(tail, full[:, k:], floor=1e-9) < tol, (flags, relerr(tail, full[:, k:], floor=1e-9))
    bad = full[:, k - 1].copy()
    bad[1, 0] *= 1.001                                   # H2 alone moved: not reachable through the three reactions
    tail, st3, _ = sim.solve_n2("pbr_k4b_solve_n2", par, zNo, tNo, period, iv=bad, k_first=k)
    assert st3[0] == 0 and st3[1] == 5 and np.isnan(tail[1]).all() and np.isfinite(tail[0]).all()


def test_forward_sensitivities_vs_central_differences():
    """K2's sensitivity integration (pbr_n1_solve<NS>, staggered direct method with the state's factorised
    Newton matrix; d r / d par from the rate compiler) on the host build: d(scaled outlet)/d(theta_k) of 2 reactors x 6 kinetic
    constants against central differences of converged plain solves.  The error scales with the integration tolerance."""
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    m = engine.Model(mi, params=tuple(theta), device=None)
    hs = HostSim(m)
    rows = cases.ensemble_rows(2)
    tab = m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})
    N, p = tab.shape[0], len(theta)
    J = np.zeros((p, 8, N))
    rel = 1e-5
    for k in range(p):
        tp, tm = tab.copy(), tab.copy()
        c = 9 + m.n + k
        tp[:, c] *= 1 + rel
        tm[:, c] *= 1 - rel
        up = hs.solve_n1(hs.prolog(tp), rtol=1e-13, atol=1e-17, n_out=2)[0][1]
        um = hs.solve_n1(hs.prolog(tm), rtol=1e-13, atol=1e-17, n_out=2)[0][1]
        J[k] = (up - um)/(2*rel*tab[:, c])[None, :]
    par = hs.prolog(np.concatenate([tab]*p))
    kp = np.repeat(np.arange(p), N)
    plain = hs.solve_n1(hs.prolog(tab), rtol=1e-9, atol=1e-13, n_out=2)
    for rtol, tol in ((1e-9, 5e-5), (1e-11, 1e-5)):
        out, sens, st, stats = hs.solve_n1_sens(par, kp, rtol=rtol, atol=rtol*1e-4)
        assert (st == 0).all()
        S = sens.reshape(8, p, N).transpose(1, 0, 2)
        for k in range(p):
            scale = np.abs(J[k]).max()
            if scale < 1e-9:                 # K3's pre-exponential: the outlet sits at that reaction's equilibrium, d/dK3 ~ 1e-12
                assert np.abs(S[k]).max() < 1e-9
                continue
            assert np.abs(S[k] - J[k]).max() < tol*scale, (rtol, k, np.abs(S[k] - J[k]).max()/scale)
        if rtol == 1e-9:                     # the sensitivities take no part in the error test: same steps, same outlet
            assert np.array_equal(stats[0][:N], plain[2][0]) and np.allclose(out[:, :N], plain[0][1], rtol=1e-11, atol=0)


def test_m7_dimensional_model_vs_reference_golden(golden):
    """the dimensional steady-state model "M7" (`rmtCore.py:300-309` -> `pbReactor.py` modelEquationM3 `:1371-1577`)
    as a variant of the scaled N1 evaluators: RHS against 200 reference states, analytic Jacobians (full and extent coordinates)
    against central differences, and the BDF solve against the reference's converged rmtExe("M7") profile."""
    g = golden("m7")
    mi = cases.m7_case()
    m = engine.Model(mi, device=None)
    assert m.variant == "M7" and m.k_in == 9 + 6 + 2
    hs = HostSim(m)
    Y, F = g["rhs_y"], g["rhs_f"]                       # dimensional [C_i, T, P]
    K = len(Y)
    par = hs.prolog(m.instance_table(mi, N=K))
    Cmax, T0, P0, zf = par[2], par[0], par[1], par[6]
    y = np.concatenate([Y[:, :6].T/Cmax, (Y[:, 7]/P0)[None], ((Y[:, 6] - T0)/T0)[None]])       # scaled [c, p, theta]
    f = hs.rhs_n1(par, y)
    fd = np.concatenate([(f[:6]*Cmax/zf).T, (f[7]*T0/zf)[:, None], (f[6]*P0/zf)[:, None]], axis=1)
    assert relerr(fd, F, floor=1e-300) < 1e-11
    # Jacobians of the variant's formulas
    f2, J = hs.jac_n1(par, y)
    for k in range(8):
        h = 1e-6*np.maximum(np.abs(y[k]), 1e-3)
        yp, ym = y.copy(), y.copy()
        yp[k] += h
        ym[k] -= h
        col = (hs.rhs_n1(par, yp) - hs.rhs_n1(par, ym))/(2*h)
        scale = np.abs(J).max(axis=(0, 1)) + 1e-300
        assert np.max(np.abs(J[:, k] - col)/scale) < 1e-6, k
    # solve: README case through M7, 30 output points (solverSetting['M9']['zNo'])
    par1 = hs.prolog(m.instance_table(mi))
    prof, st, stats = hs.solve_n1(par1, rtol=1e-9, atol=1e-13, n_out=30)
    assert st[0] == 0
    C = prof[:, :6, 0]*par1[2]
    data = np.concatenate([(C/C.sum(axis=1, keepdims=True)).T, (prof[:, 7, 0]*par1[0] + par1[0])[None]])
    assert relerr(data, g["conv_dataYs"]) < 1e-6
    assert relerr(data, g["default_dataYs"]) < 2e-3                  # the reference's own default-tolerance run
    ens = O.Ensemble(mi)
    assert relerr(ens.rhs_m7(Y), F, floor=1e-300) < 1e-12           # the oracle restatement against the same reference states
    # extent coordinates (what K2 integrates): f = M f_x and M J_x = J M for the variant's formulas as well
    nx, R = hs.L.hs_nx(), m.n_rxn
    rows = cases.ensemble_rows(8)
    par8 = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    rng = np.random.default_rng(1)
    x = np.zeros((nx, 8))
    x[:R] = rng.uniform(0, 0.02, (R, 8)); x[R] = rng.uniform(0.8, 1, 8); x[R + 1] = rng.uniform(0, 0.1, 8)
    yx, fx, Jx = hs.n1x(par8, x)
    ff, JJ = hs.jac_n1(par8, yx)
    M = _extent_map(m, hs.nv1, nx)
    assert relerr(M @ fx, ff, floor=1e-300) < 1e-12
    for i in range(8):
        assert np.max(np.abs(M @ Jx[:, :, i] - JJ[:, :, i] @ M)) < 1e-12*np.max(np.abs(JJ[:, :, i]))


def test_block_kernels_do_not_depend_on_the_thread_schedule():
    """Race check (compute-sanitizer is closed on the GPU pool): the emulator resumes the threads of a block in ascending,
    descending and per-round random order.  A shared-memory race -- a read and a write of the same location with no barrier or
    warp synchronisation in between -- makes the result depend on that order; every block kernel of the N2 path (serial march,
    scan march, slab history, on-chip 512-thread variant, resume projection) and K3 (block scan with double-buffered totals, one
    barrier per round) must give bit-identical results under all three."""
    mi, m, hs, sim = _simt_case(lambda: cases.readme_case("N2"))
    rows = cases.ensemble_rows(3)
    par = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    zNo, tNo, period = 40, 2, 0.004
    rng = np.random.default_rng(2)
    Y = np.zeros((3, m.nv2, 45))
    Y[:, :6] = par[16:22].T[:, :, None]*np.exp(rng.normal(0, 0.2, (3, 6, 45)))
    Y[:, 6] = rng.uniform(0, 0.2, (3, 45))
    ref_rhs = hs.rhs_n2(par, Y)
    base = {}
    try:
        # negative control: the check must see a real race (a neighbour exchange without a barrier) and pass its fixed version
        demo = {}
        for mode in (0, 1, 2):
            sim.set_order(mode, seed=99)
            demo[mode] = (sim.race_demo(True), sim.race_demo(False))
        assert not np.array_equal(demo[0][0], demo[1][0]) and not np.array_equal(demo[0][0], demo[2][0])
        assert np.array_equal(demo[0][1], demo[1][1]) and np.array_equal(demo[0][1], demo[2][1])
        for mode in (0, 1, 2):
            sim.set_order(mode, seed=12345 + mode)
            for kernel in ("pbr_k4b_solve_n2", "pbr_k4bs_solve_n2", "pbr_k4c_solve_n2", "pbr_k4d_solve_n2"):
                if mode == 1 and kernel in ("pbr_k4c_solve_n2", "pbr_k4d_solve_n2"):
                    continue                                   # same body as the first two (template flags); modes 0 and 2 cover them
                snaps, status, stats = sim.solve_n2(kernel, par[:, :2], zNo, tNo, period, grid=1)
                assert (status == 0).all()
                if mode == 0:
                    base[kernel] = (snaps, stats)
                else:
                    assert np.array_equal(snaps, base[kernel][0]) and np.array_equal(stats, base[kernel][1]), (kernel, mode)
            tail, st, _ = sim.solve_n2("pbr_k4b_solve_n2", par[:, :2], zNo, tNo, period, iv=base["pbr_k4b_solve_n2"][0][:, 0], k_first=1)
            if mode == 0:
                base["resume"] = tail
            else:
                assert np.array_equal(tail, base["resume"]), mode
            for G, staged in ((1, 1), (3, 1), (3, 0)):         # 45 nodes per reactor: ragged rounds, several reactors per block
                dY = sim.rhs_n2(par, Y, G, staged)
                assert relerr(dY, ref_rhs, floor=1e-12) < 1e-12, (mode, G, staged)
                if mode == 0:
                    base[("k3", G, staged)] = dY
                else:
                    assert np.array_equal(dY, base[("k3", G, staged)]), (mode, G, staged)
    finally:
        sim.set_order(0)


def test_m8_dimensional_model_with_velocity_vs_reference_golden(golden):
    """model "M8" (`rmtCore.py:311-320` -> `pbReactor.py` modelEquationM4 `:1778-1995`; unknowns C_i, T, P, v;
    the reference's built-in kinetics) in the scaled flux variables: RHS against 200 reference states, f = M f_x between full and
    extent coordinates, the forward-difference Jacobian against central differences, and the BDF solve against the reference's
    converged rmtExe("M8") series [x_i ; T ; P ; v]."""
    from pyremot_b200 import api
    g = golden("m8")
    mi = api._m8_input(cases.m8_case())
    m = engine.Model(mi, device=None)
    assert m.variant == "M8" and m.k_in == 9 + 6 + 3 and m.nv1 == 9 and m.n1_vars == 6
    hs = HostSim(m)
    Y, F = g["rhs_y"], g["rhs_f"]                        # dimensional [C_i, T, P, v]
    K = len(Y)
    par = hs.prolog(m.instance_table(mi, N=K))
    Cmax, T0, P0, zf, v0 = par[2], par[0], par[1], par[6], par[4]
    C, T, P, v = Y[:, :6].T, Y[:, 6], Y[:, 7], Y[:, 8]
    y = np.concatenate([C*v/(Cmax*v0), (P/P0)[None], ((T - T0)/T0)[None], (v/v0)[None]])        # [phi, p, theta, u]
    f = hs.rhs_n1(par, y)
    dT, dP, dv = f[7]*T0/zf, f[6]*P0/zf, f[8]*v0/zf
    dC = (f[:6]*Cmax*v0/zf - C*dv)/v                     # d(C v)/dz = r  ->  dC/dz = (r - C dv/dz)/v, the reference's form (:1969-1971)
    fd = np.concatenate([dC.T, dT[:, None], dP[:, None], dv[:, None]], axis=1)
    assert relerr(fd, F, floor=1e-300) < 1e-10
    # extent coordinates and the forward-difference Jacobian
    nx, R = hs.L.hs_nx(), m.n_rxn
    rows = cases.ensemble_rows(6)
    A = np.pi*0.0381**2/4
    par6 = hs.prolog(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}))
    rng = np.random.default_rng(4)
    x = np.zeros((nx, 6))
    x[:R] = rng.uniform(0, 0.02, (R, 6)); x[R] = rng.uniform(0.9, 1, 6); x[R + 1] = rng.uniform(0, 0.1, 6); x[R + 2] = rng.uniform(0.9, 1.1, 6)
    yx, fx, Jx = hs.n1x(par6, x)
    ff = hs.rhs_n1(par6, yx)
    Mx = np.zeros((9, nx))
    Mx[:6, :R] = m.S.T; Mx[6, R] = Mx[7, R + 1] = Mx[8, R + 2] = 1.0
    assert relerr(Mx @ fx, ff, floor=1e-300) < 1e-12
    for k in range(nx):
        h = 1e-6*np.maximum(np.abs(x[k]), 1e-2)
        xp, xm = x.copy(), x.copy()
        xp[k] += h; xm[k] -= h
        col = (hs.n1x(par6, xp)[1] - hs.n1x(par6, xm)[1])/(2*h)
        assert np.max(np.abs(Jx[:, k] - col)/(np.abs(Jx).max(axis=(0, 1)) + 1e-300)) < 1e-5, k
    # solve: README inputs through M8, 25 output points (solverSetting['S3']['timesNo'])
    par1 = hs.prolog(m.instance_table(mi))
    prof, st, stats = hs.solve_n1(par1, rtol=1e-9, atol=1e-13, n_out=25)
    assert st[0] == 0
    phi = prof[:, :6, 0]
    series = np.concatenate([(phi/phi.sum(axis=1, keepdims=True)).T, (prof[:, 7, 0]*par1[0] + par1[0])[None], (prof[:, 6, 0]*par1[1])[None],
                             (prof[:, 8, 0]*par1[4])[None]])
    assert relerr(series, g["conv_series"]) < 1e-6
    assert relerr(series[:7], g["default_dataYs"]) < 2e-3
    assert relerr(O.Ensemble(cases.m8_case()).rhs_m8(Y), F, floor=1e-300) < 1e-12


def test_m9_dynamic_dimensional_model_vs_reference_golden(golden):
    """model "M9" (`rmtCore.py:322-331` -> `pbReactor.py` modelEquationM5 `:2296-2666`) through the warp
    solver's own sweep: RHS against the reference on random states at two grids (pressure AND velocity marched through the node
    evaluations, one species below the clamp), the analytic diagonal blocks against central differences of the node
    function, and the BDF interval loop against the reference's converged rmtExe("M9") temperature profiles."""
    g = golden("m9")
    mi = cases.m9_case()
    m = engine.Model(mi, device=None)
    assert m.variant == "M9" and m.k_in == 9 + 6 + 5 and m.nv2 == 7
    hs = HostSim(m)
    for zNo in (12, 40):
        Y, F = g[f"rhs_y_z{zNo}"], g[f"rhs_f_z{zNo}"]                      # dimensional [C_i ; T] per node
        par = hs.prolog(m.instance_table(mi, N=len(Y)))
        Cmax, T0 = par[2][:, None, None], par[0][:, None]
        y = Y.copy()
        y[:, :6] /= Cmax
        y[:, 6] = (Y[:, 6] - T0)/T0
        f, Jd, Js = hs.rhs_n2(par, y, jac=True)
        fd = f.copy()
        fd[:, :6] *= Cmax
        fd[:, 6] *= T0
        assert np.max(np.abs(fd - F)/(np.abs(F) + 1e-12*np.abs(F).max())) < 1e-11, zNo
        # the diagonal blocks: perturb ONE node's state; the marched (P, v) at that node depend on upstream nodes only, so a central
        # difference of the full RHS restricted to the node's own rows is the block the solver builds analytically
        i, z = 0, zNo//3
        par1 = np.ascontiguousarray(par[:, i:i + 1])
        for k in range(7):
            h = 1e-4*max(abs(y[i, k, z]), 1e-10)            # relative step: the trace species sit at 1e-6 and the rates go like 1/c there
            yp, ym = y[i:i + 1].copy(), y[i:i + 1].copy()
            yp[0, k, z] += h
            ym[0, k, z] -= h
            col = (hs.rhs_n2(par1, yp)[0, :, z] - hs.rhs_n2(par1, ym)[0, :, z])/(2*h)
            assert np.max(np.abs(Jd[i, z, :, k] - col))/np.abs(Jd[i, z]).max() < 1e-7, (zNo, k)    # analytic blocks vs central differences
    for tag, mic, zNo in (("z12", cases.m9_case(period=0.2), 12), ("z20", cases.m9_case(period=0.1, T=540.0, Tm=530.0), 20)):
        par = hs.prolog(m.instance_table(mic))
        snaps, st, stats = hs.solve_n2(par, zNo, 2, float(mic["operating-conditions"]["period"]), rtol=1e-9, atol=1e-13)
        assert st[0] == 0
        T = snaps[0, :, 6, :]*par[0, 0] + par[0, 0]
        assert np.max(np.abs(T - g[f"{tag}_conv_T"]))/500 < 1e-10, tag
        assert np.max(np.abs(T - g[f"{tag}_default_T"]))/500 < 1e-6, tag        # the reference's own LSODA run at rtol 1e-3
        # sub-diagonal d f_z / d y_{z-1} on the solved (physical) state: at frozen (P_z, v_z) it is diagonal; what the Newton matrix
        # leaves out -- node z-1 also moves (P_z, v_z) through the march -- is a percent-level effect there
        ys = np.ascontiguousarray(snaps[:, -1])
        _, Jd, Js = hs.rhs_n2(par, ys, jac=True)
        z = zNo//2
        for k in range(7):
            h = 1e-6*max(abs(ys[0, k, z - 1]), 1e-2)
            yp, ym = ys.copy(), ys.copy()
            yp[0, k, z - 1] += h
            ym[0, k, z - 1] -= h
            col = (hs.rhs_n2(par, yp)[0, :, z] - hs.rhs_n2(par, ym)[0, :, z])/(2*h)
            assert abs(Js[0, z, k] - col[k]) < 2e-2*abs(col[k]), (tag, k, Js[0, z, k], col[k])


def test_m9_on_the_block_per_reactor_kernels(golden):
    """Model M9 through the block-per-reactor N2 solver on the SIMT emulator: the march of pressure and velocity is solved by
    block-wide fixed-point sweeps (pbr_k4b_eval), full coordinates (7 variables per node).  One warp (12 nodes) against the
    reference's converged rmtExe("M9") profiles; two warps (40 nodes), serial and scan march, against the scalar warp solver."""
    from hostsim import SimtSim
    g = golden("m9")
    mi = cases.m9_case(period=0.2)
    m = engine.Model(mi, device=None)
    hs, sim = HostSim(m), SimtSim(m)
    par = hs.prolog(m.instance_table(mi))
    snaps, st, stats = sim.solve_n2("pbr_k4b_solve_n2", par, 12, 2, 0.2, rtol=1e-9, atol=1e-13)
    assert st[0] == 0
    T = snaps[0, :, 6, :]*par[0, 0] + par[0, 0]
    assert np.max(np.abs(T - g["z12_conv_T"]))/500 < 1e-10
    ref, st0, stats0 = hs.solve_n2(par, 40, 2, 0.2, rtol=1e-9, atol=1e-13)
    for kernel in ("pbr_k4b_solve_n2", "pbr_k4bs_solve_n2"):
        snaps, st, stats = sim.solve_n2(kernel, par, 40, 2, 0.2, rtol=1e-9, atol=1e-13)
        assert st[0] == 0 and stats[0, 0] == stats0[0, 0]
        assert np.max(np.abs(snaps - ref)) < 1e-12, kernel
