"""GPU parity of the fused N2 RHS kernel (K3: one warp per reactor, lanes over axial nodes) against the
reference's modelEquationN2 outputs (tests/golden/n2_rhs.npz) and the oracle at the BASELINE node counts."""
import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases

pytestmark = pytest.mark.gpu


def relerr(a, b, floor=0.0):
    return float(np.max(np.abs(np.asarray(a) - b)/(np.abs(b) + floor)))


@pytest.mark.parametrize("tag,mk", [
    ("readme_z5", lambda: cases.readme_case("N2")),
    ("readme_z12", lambda: cases.readme_case("N2")),
    ("readme_iso_z8", lambda: cases.readme_case("N2", "iso-thermal")),
    ("dme_z20", cases.dme_n2_case),
    ("ch4_z10", cases.ch4_n2_case),
])
def test_rhs_n2_vs_reference_golden(golden, tag, mk):
    import torch
    from pyremot_b200.api import get_model
    g = golden("n2_rhs")
    Y, F = g[f"y_{tag}"], g[f"f_{tag}"]
    mi = mk()
    m = get_model(mi)
    par = m.prolog(torch.from_numpy(m.instance_table(mi, N=len(Y))).cuda())
    f = m.rhs_n2(par, torch.from_numpy(np.ascontiguousarray(Y)).cuda()).cpu().numpy()
    assert relerr(f, F, floor=1e-9) < 1e-11


@pytest.mark.parametrize("zNo,N", [(100, 512), (400, 64), (33, 7), (2, 3)])
def test_rhs_n2_vs_oracle_at_baseline_grids(zNo, N):
    """zNo=100 / 400 are the BASELINE config 3 / 4 grids; 33 and 2 exercise ragged warps and the minimum grid."""
    import torch
    from pyremot_b200.api import get_model
    rows = cases.ensemble_rows(N)
    sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}
    mi = cases.readme_case("N2")
    m = get_model(mi)
    ens = O.Ensemble(mi, sweep)
    rng = np.random.default_rng(zNo)
    Y = ens.iv_n2(zNo)
    Y[:, :6] *= np.exp(rng.normal(0, 0.3, Y[:, :6].shape))
    Y[:, 6] = rng.uniform(0, 0.25, Y[:, 6].shape)
    Y[::5, 2, zNo//2] = -1e-7            # below the 1e-30 clamp
    par = m.prolog(torch.from_numpy(m.instance_table(mi, sweep)).cuda())
    f = m.rhs_n2(par, torch.from_numpy(Y).cuda()).cpu().numpy()
    # random states include nodes where the two terms of a reversible rate nearly cancel: the few-ulp differences between the
    # compiled DAG (1/exp(u) -> exp(-u), shared reciprocals) and NumPy's evaluation order are amplified there
    assert relerr(f, ens.rhs_n2(Y), floor=1e-9) < 3e-10
