"""Short single-GPU program for ncu: one K0+K2+K5 pass over N reactors and a few K3 launches.

    python tools/prof_run.py [N_n1] [N_n2] [zNo]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N1 = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N2 = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
zNo = int(sys.argv[3]) if len(sys.argv) > 3 else 100

mi = cases.readme_case("N1")
m = get_model(mi)
rows = cases.ensemble_rows(N1)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
for _ in range(2):
    par = m.prolog(tab)
    r = m.solve_n1(par, rtol=1e-9, atol=1e-13, n_out=101, profile=False)
    out = m.unscale_n1(par, r["outlet_y"][None])
torch.cuda.synchronize()
st = r["stats"].cpu().numpy()
print("N1", N1, "status!=0:", int((r["status"] != 0).sum()), "mean steps/nfev/nlu", st[0].mean(), st[2].mean(), st[3].mean())

if N2 > 0:
    mi2 = cases.readme_case("N2")
    m2 = get_model(mi2)
    rows = cases.ensemble_rows(N2)
    tab2 = torch.from_numpy(m2.instance_table(mi2, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
    par2 = m2.prolog(tab2)
    Y = torch.empty((N2, m2.nv2, zNo), dtype=torch.float64, device="cuda")
    Y[:, :m2.n, :] = par2[16:16 + m2.n].t()[:, :, None]
    Y[:, m2.n, :] = torch.linspace(0, 0.2, zNo, device="cuda", dtype=torch.float64)[None, :]
    dY = torch.empty_like(Y)
    for _ in range(3):
        m2.rhs_n2(par2, Y, out=dY)
    torch.cuda.synchronize()
    print("N2 rhs", N2, zNo, float(dY.abs().max()))

if os.environ.get("PBR_SAVE_CUBIN"):      # exact cubin of this run, for tools/sass_profile.py
    import ctypes as _C
    from pyremot_b200 import engine as _E
    _d, _s = _C.c_void_p(), _C.c_ulong()
    _E.lib().pbr_model_cubin(m._ctx, _C.byref(_d), _C.byref(_s))
    open(os.environ["PBR_SAVE_CUBIN"], "wb").write(_C.string_at(_d.value, _s.value))
    open(os.environ["PBR_SAVE_CUBIN"] + ".cu", "w").write(m.source)
