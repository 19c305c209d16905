"""M9 ensemble (1,184 reactors, 100 nodes, 10 intervals, 2 s) on the warp-per-reactor kernel with 2 ... 8 persistent warps per SM:
the per-warp scratch slab is 152 KB at this grid, so 8 warps per SM (180 MB) spill past the 126 MB L2, 4 (90 MB) do not.

    python tools/m9_warps_ab.py > gpurun_out/r2_m9_warps_ab.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N, zNo, tNo, period = 1184, 100, 10, 2.0
mi = cases.m9_case(period=period)
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"] - 10.0})).cuda()
par = m.prolog(tab)
out = {}
for w in (8, 6, 4, 3, 2):
    os.environ["PBR_N2_WARPS_PER_SM"] = str(w)
    ts = []
    for k in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m.solve_n2(par, zNo, tNo, period)
        e1.record()
        torch.cuda.synchronize()
        if k:
            ts.append(e0.elapsed_time(e1))
    out[f"warps_{w}"] = {"ms": min(ts), "failed": int((r["status"] != 0).sum().item())}
print(json.dumps(out))
