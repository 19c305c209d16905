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
