"""C-ABI library on a machine without a GPU: it loads, exports every symbol include/pbr.h declares, compiles
models through NVRTC for sm_100a (compile-only context) and refuses compute calls loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pyremot_b200 import cases, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pbr.h")).read()
    declared = set(re.findall(r"^(?:int|long|void|const char\*|pbr_ctx\*)\s+(pbr_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(engine.EXPORTS)
    L = C.CDLL(engine.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in engine.lib().pbr_version()


def test_nvrtc_compile_only_for_sm_100a(tmp_path):
    m = engine.Model(cases.readme_case("N1"), device=None, cache_dir=str(tmp_path))
    assert (m.nv1, m.nv2, m.k_in, m.k_par, m.n_rxn, m.n_par) == (8, 7, 15, 22, 3, 0)
    src = m.source
    for k in ("pbr_k0_prolog", "pbr_k2_solve_n1", "pbr_k3_rhs_n2", "pbr_k4_solve_n2", "pbr_k5_unscale_n1", "pbr_rates_jac",
              "#define PBR_NC 6", "#define PBR_S_NNZ(F) F(0,0,-3.0)"):
        assert k in src, k
    data, size = C.c_void_p(), C.c_ulong()
    assert engine.lib().pbr_model_cubin(m._ctx, C.byref(data), C.byref(size)) == 0
    blob = C.string_at(data.value, size.value)
    assert blob[:4] == b"\x7fELF" and len(os.listdir(tmp_path)) == 1          # cubin cached on disk
    m2 = engine.Model(cases.readme_case("N1"), device=None, cache_dir=str(tmp_path))   # second load hits the cache
    assert m2.source == src
    iso = engine.Model(cases.readme_case("N1", "iso-thermal"), device=None, cache_dir=str(tmp_path))
    assert (iso.nv1, iso.nv2) == (7, 6)
    ch4 = engine.Model(cases.ch4_n2_case(), device=None, cache_dir=str(tmp_path))
    assert (ch4.nv1, ch4.nv2, ch4.n_rxn) == (5, 4, 1)


def test_compute_calls_fail_loudly_without_a_device(tmp_path):
    m = engine.Model(cases.readme_case("N1"), device=None, cache_dir=str(tmp_path))
    tab = m.instance_table(cases.readme_case("N1"))
    with pytest.raises(engine.PbrError, match="compile-only"):
        m.solve_n1_host(tab)
    with pytest.raises(engine.PbrError, match="compile-only"):
        m.solve_n2_host(tab, 20, 5, 0.5)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(engine.PbrError, match="no usable CUDA device"):
            engine.Model(cases.readme_case("N1"), device=0, cache_dir=str(tmp_path))
        from pyremot_b200 import rmtExe
        with pytest.raises(engine.PbrError):
            rmtExe(cases.readme_case("N1"))


def test_bad_rate_source_reports_nvrtc_log():
    L = engine.lib()
    ctx = L.pbr_create(-1)
    one = (C.c_double*4)(1, 0, 0, 0)
    assert L.pbr_load_rates(ctx, b"", 0, b"", None) == -1                      # species not set yet
    assert L.pbr_set_species(ctx, 1, 1, one, one, one, one, one, one, 0) == 0
    log = C.c_char_p()
    rc = L.pbr_load_rates(ctx, b"#define PBR_NP 0\nthis is not cuda;", 0, b"", C.byref(log))
    assert rc == -3 and b"error" in log.value and b"NVRTC" in L.pbr_last_error(ctx)
    assert L.pbr_set_species(ctx, 0, 1, one, one, one, one, one, one, 0) == -1
    L.pbr_destroy(ctx)


def test_instance_table_and_reaction_tables():
    mi = cases.readme_case("N1", theta_as_vars=True)
    m = engine.Model(mi, params=("A_K1", "A_KCO"), device=None)
    t = m.instance_table(mi)
    assert t.shape == (1, 17)
    assert list(t[0, :9]) == [523.0, 5e6, 0.000228, 0.0381, 1, 0.002, 0.39, 50.0, 523.0]
    assert list(t[0, 9:15]) == mi["feed"]["concentration"] and list(t[0, 15:]) == [35.45, 7.99e-7]
    t = m.instance_table(mi, {"T": np.array([500.0, 510.0]), "VARS:A_KCO": np.array([1e-6, 2e-6])})
    assert t.shape == (2, 17) and list(t[:, 0]) == [500.0, 510.0] and list(t[:, 16]) == [1e-6, 2e-6] and list(t[:, 1]) == [5e6, 5e6]
    with pytest.raises(KeyError):
        m.instance_table(mi, {"VARS:A_K2": np.array([1.0])})
    S, dH25, dcp = engine.parse_reactions(cases.MEOH_DME_REACTIONS, cases.MEOH_DME_COMPONENTS)
    assert np.array_equal(S, [[-3, -1, 1, 0, 1, 0], [1, 1, -1, -1, 0, 0], [0, 0, 1, 0, -2, 1]])
    assert np.allclose(dH25, [-49010, -41160, -24520], rtol=1e-12)             # tests/test_heatCapacity.py values
    # a species that is not in the component list drops out of S but still counts in dH25 / dCp
    S2, dH2, _ = engine.parse_reactions({"R": "CO + H2O <=> H2 + CO2"}, ["H2", "CO2", "H2O"])
    assert np.array_equal(S2, [[1, 1, -1]]) and np.isclose(dH2[0], -41160)
    from pyremot_b200 import rmtCom
    assert rmtCom() == "CO2,H2,CH3OH,H2O,CO,DME,N2,CH4,C2H4,C3H6,C3H8,C4H10"


def test_perturbation_table_for_parameter_estimation():
    from pyremot_b200 import estimation
    mi = cases.readme_case("N1", theta_as_vars=True)
    theta = list(cases.MEOH_DME_THETA)
    m = engine.Model(mi, params=theta, device=None)
    rows = cases.ensemble_rows(5)
    tab, steps, N = estimation.perturbation_table(m, mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}, theta, 1e-4)
    assert tab.shape == (5*7, 21) and N == 5 and steps.shape == (5, 6)
    assert np.array_equal(tab[:5], tab[5:10][:, :].copy()*0 + tab[:5])            # base block untouched
    for k in range(6):
        blk = tab[5*(k+1):5*(k+2)]
        diff = blk - tab[:5]
        assert np.count_nonzero(diff) == 5 and np.allclose(diff[:, 15 + k], tab[:5, 15 + k]*1e-4, rtol=1e-12)
