"""Device elementary functions used by the compiled rate expressions (csrc/pbr_math.cuh: branch-free rcp / log / sqrt,
libdevice exp) checked directly on the GPU: a three-"reaction" model whose rates ARE those functions, read back through
the iso-thermal N1 RHS at the inlet state (there w = 1 and f_i = r_i / G, so r_i = f_i * G exactly up to one rounding)."""
import math

import numpy as np
import pytest

from pyremot_b200 import cases

pytestmark = pytest.mark.gpu


def _model_input(rates):
    mi = cases.readme_case("N1", "iso-thermal")
    mi["feed"]["components"]["shell"] = ["H2", "CO2", "H2O"]
    mi["feed"]["concentration"] = [10.0, 20.0, 30.0]
    mi["reactions"] = {"R1": " <=> H2", "R2": " <=> CO2", "R3": " <=> H2O"}     # each reaction feeds one species
    mi["reaction-rates"] = {"VARS": {}, "RATES": rates}
    return mi


@pytest.mark.parametrize("name,rates,ref", [
    ("log/log10/sqrt", {"r1": lambda x: math.log(x["T"]), "r2": lambda x: math.log10(x["P"]), "r3": lambda x: math.sqrt(x["P"]*1e-5)},
     lambda T, P: [np.log(T), np.log10(P), np.sqrt(P*1e-5)]),
    ("exp/rcp/pow", {"r1": lambda x: math.exp(-17069.0/(8.314472*x["T"])), "r2": lambda x: 1/x["T"] + 3.0/x["P"],
                     "r3": lambda x: math.pow(x["T"], -1.5) + x["T"]**3},
     lambda T, P: [np.exp(-17069.0/(8.314472*T)), 1/T + 3.0/P, T**-1.5 + T**3]),
    ("small and large arguments", {"r1": lambda x: math.log(x["T"]*1e-300) + math.log(x["P"]*1e200),
                                   "r2": lambda x: math.sqrt(x["T"]*1e-250)*math.sqrt(x["P"]*1e250),
                                   "r3": lambda x: math.exp(x["T"]*-1.2) + math.exp(x["T"]/100.0)},
     lambda T, P: [np.log(T*1e-300) + np.log(P*1e200), np.sqrt(T*1e-250)*np.sqrt(P*1e250), np.exp(T*-1.2) + np.exp(T/100.0)]),
])
def test_device_math_matches_numpy(name, rates, ref):
    import torch
    from pyremot_b200.api import get_model
    mi = _model_input(rates)
    m = get_model(mi)
    N = 4096
    rng = np.random.default_rng(11)
    T = 300 + 700*rng.random(N)
    P = 10.0**rng.uniform(4, 7.5, N)
    tab = m.instance_table(mi, {"T": T, "P": P})
    par = m.prolog(torch.from_numpy(tab).cuda())
    y = torch.cat([par[16:19], torch.ones((1, N), dtype=torch.float64, device="cuda")]).contiguous()   # inlet state
    f = m.rhs_n1(par, y).cpu().numpy()
    G = par[11].cpu().numpy()
    r = f[:3]*G
    want = np.array(ref(T, P))
    err = np.max(np.abs(r - want)/np.abs(want))
    assert err < 2e-15*8, (name, err)          # a few ulp: function (<= 1.5 ulp) + the f*G round trip


def test_reduced_precision_primitives_stay_inside_their_bounds():
    """The integrators steer with reduced-precision reciprocals (pbr_rcp_20bit: error-norm weights; pbr_rcp_40bit: LU pivots and
    Jacobian-only factors) and a single-precision x^e for step-size factors.  The host emulation replaces them by libm, so their
    error bounds are checked here on the device, over the magnitudes they see (1e-300 .. 1e300 for the reciprocals)."""
    import torch
    from pyremot_b200.api import get_model
    m = get_model(cases.readme_case("N1"))
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0**rng.uniform(-290, 290, 200000), 1.0 + rng.random(100000), 10.0**rng.uniform(-13, 3, 200000)])
    x *= np.where(rng.random(x.size) < 0.3, -1.0, 1.0)
    out = m.math_probe(torch.from_numpy(x).cuda()).cpu().numpy()
    inv = 1.0/x
    e_full, e40, e20 = (np.abs(out[k]*x - 1.0).max() for k in (0, 1, 2))
    print(f"\nrcp: full {e_full:.2e}, 40-bit {e40:.2e}, 20-bit {e20:.2e}")
    assert e_full < 4e-16 and e40 < 1e-11 and e20 < 2e-6          # <= 1 ulp, one Newton step of 2^-20, the bare approximation
    assert (np.sign(out[2]) == np.sign(inv)).all()
    pos = np.abs(x[x.size//2:])
    outp = m.math_probe(torch.from_numpy(pos).cuda()).cpu().numpy()
    assert (np.abs(outp[3] - np.log(pos))/np.maximum(np.abs(np.log(pos)), 1.0)).max() < 4e-16
    assert (np.abs(outp[4] - np.sqrt(pos))/np.sqrt(pos)).max() < 3e-16
    small = pos[pos < 700]
    outs = m.math_probe(torch.from_numpy(small).cuda()).cpu().numpy()
    assert (np.abs(outs[5] - np.exp(-small))/np.exp(-small)).max() < 5e-16
    # step-size factors: x^(-1/3) to 3 digits is enough (they are clipped to [0.2, 10] and multiplied by a 0.9 safety factor)
    f = pos[(pos > 1e-12) & (pos < 1e12)]
    outf = m.math_probe(torch.from_numpy(f).cuda()).cpu().numpy()
    assert (np.abs(outf[6] - f**(-1.0/3.0))/f**(-1.0/3.0)).max() < 1e-5
