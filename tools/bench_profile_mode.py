"""K2 with and without the 101-point dense output (rmtExe's default asks for the profile)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mi = cases.readme_case("N1")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m.prolog(tab)
for prof in (False, True):
    for _ in range(2):
        r = m.solve_n1(par, profile=prof)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = m.solve_n1(par, profile=prof)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    extra = ""
    if prof:
        e0.record()
        out = m.unscale_n1(par, r["prof_y"])
        e1.record()
        torch.cuda.synchronize()
        extra = f"  K5 un-scale of the profile: {e0.elapsed_time(e1):.2f} ms ({out.numel()*8/1e6:.0f} MB)"
    print(f"profile={prof}: K2 {ms:.2f} ms for {N} solves{extra}")
