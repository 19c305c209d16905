"""Short N2 solve for ncu: python tools/prof_n2.py [N] [zNo] [tNo] [period]      (PBR_PROF_MODEL=M9: the dynamic dimensional model)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
zNo = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tNo = int(sys.argv[3]) if len(sys.argv) > 3 else 1
period = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
m9 = os.environ.get("PBR_PROF_MODEL") == "M9"
mi = cases.m9_case(period=period) if m9 else cases.readme_case("N2")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - (10.0 if m9 else 0.0)})).cuda()
par = m.prolog(tab)
for _ in range(2):
    r = m.solve_n2(par, zNo, tNo, period)
torch.cuda.synchronize()
st = r["stats"].cpu().numpy()
print("N2", N, zNo, "failed", int((r["status"] != 0).sum()), "mean steps", st[0].mean(), "nfev", st[2].mean())

if os.environ.get("PBR_SAVE_CUBIN"):      # exact cubin of this run, for tools/sass_profile.py
    import ctypes as _C
    from pyremot_b200 import engine as _E
    _d, _s = _C.c_void_p(), _C.c_ulong()
    _E.lib().pbr_model_cubin(m._ctx, _C.byref(_d), _C.byref(_s))
    open(os.environ["PBR_SAVE_CUBIN"], "wb").write(_C.string_at(_d.value, _s.value))
    open(os.environ["PBR_SAVE_CUBIN"] + ".cu", "w").write(m.source)
