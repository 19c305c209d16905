"""Small single-GPU program for compute-sanitizer (memcheck / racecheck / initcheck / synccheck, ONE tool per gpurun call):
every kernel family of the path once, at sizes that finish in seconds under the tool.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402


def table(m, mi, N, offset=0):
    rows = cases.ensemble_rows(N, offset=offset)
    return torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()


done = []
mi1 = cases.readme_case("N1")
m1 = get_model(mi1)
par = m1.prolog(table(m1, mi1, 300))                                   # K0; 300: a ragged last block
r = m1.solve_n1(par, n_out=11, profile=True)                           # K2 (shared-memory history + LU, block barrier per trip)
out = m1.unscale_n1(par, r["prof_y"])                                  # K5
y = torch.rand((m1.nv1, 300), dtype=torch.float64, device="cuda")*0.1 + 0.5
m1.rhs_n1(par, y); m1.jac_n1(par, y)                                   # standalone RHS / Jacobian
torch.cuda.synchronize()
assert int((r["status"] != 0).sum()) == 0
done.append("K0 K2 K5 rhs_n1 jac_n1")

mis = cases.readme_case("N1", theta_as_vars=True)
ms = get_model(mis, params=tuple(cases.MEOH_DME_THETA))
tab = table(ms, mis, 100)
kp = torch.arange(6, dtype=torch.int32, device="cuda").repeat_interleave(100)
rs = ms.solve_n1_sens(ms.prolog(tab.repeat(6, 1).contiguous()), kp)    # K2s + K5s
torch.cuda.synchronize()
assert int((rs["status"] != 0).sum()) == 0
done.append("K2s K5s")

mi2 = cases.readme_case("N2")
m2 = get_model(mi2)
for zNo in (100, 33):
    p2 = m2.prolog(table(m2, mi2, 37))
    Y = torch.rand((37, m2.nv2, zNo), dtype=torch.float64, device="cuda")*0.2 + 0.2
    m2.rhs_n2(p2, Y)                                                   # K3 (block scan, double-buffered totals)
torch.cuda.synchronize()
done.append("K3")


def n2(N, zNo, tNo, period, env=None, **kw):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        p2 = m2.prolog(table(m2, mi2, N))
        r2 = m2.solve_n2(p2, zNo, tNo, period, **kw)
        m2.unscale_n2(p2, r2["snaps_y"])                               # K5b
        torch.cuda.synchronize()
        assert int((r2["status"] != 0).sum()) == 0, (zNo, r2["status"])
        return r2
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


r2 = n2(6, 40, 2, 0.02)                                                # K4b: serial march, 2 warps (ragged)
done.append("K4b")
n2(6, 40, 3, 0.03, iv=r2["snaps_y"][:, 0].contiguous(), k_first=1)     # K4b resume path (projection on the extents)
done.append("K4b-resume")
n2(3, 160, 1, 0.01)                                                    # K4bs: scan march, 5 warps
done.append("K4bs")
n2(2, 300, 1, 0.005)                                                   # K4d: all on chip, <= 512 threads
done.append("K4d")
n2(2, 300, 1, 0.005, env={"PBR_K4C_ONCHIP": "0"})                      # K4c: history in a global slab
done.append("K4c")
n2(2, 530, 1, 0.003)                                                   # K4: warp per reactor (grids above 512 nodes)
done.append("K4")
print("sanitize_run: ok:", ", ".join(done))
