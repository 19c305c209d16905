"""Serial march (pbr_k4b_solve_n2) against the scan march (pbr_k4bs_solve_n2) of the block-per-reactor N2 solver over the grid
sizes both can run (the launcher's switch is PBR_K4B_SCAN_FROM, read at every launch): 592 reactors, tNo 5, 0.5 s; device time,
median of 3, and the worst difference between the two results.

    python tools/k4b_march_ab.py [N [zNo,zNo,...]] > gpurun_out/r2_k4b_march_ab.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

mi = cases.readme_case("N2")
m = get_model(mi)
out = {}
# optional arguments: N and a comma list of grid sizes (default: one wave of 592 / 296 reactors over all grid sizes)
N_arg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
grids = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (10, 20, 33, 64, 100, 128, 200, 256)
for zNo in grids:
    N = N_arg or (592 if zNo <= 128 else 296)
    rows = cases.ensemble_rows(N)
    par = m.prolog(torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda())
    rec = {}
    res = {}
    for name, frm in (("serial", 100000), ("scan", 0)):
        os.environ["PBR_K4B_SCAN_FROM"] = str(frm)
        ts = []
        for k in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = m.solve_n2(par, zNo, 5, 0.5)
            e1.record()
            torch.cuda.synchronize()
            if k:
                ts.append(e0.elapsed_time(e1))
        rec[name + "_ms"] = float(np.median(ts))
        rec[name + "_failed"] = int((r["status"] != 0).sum().item())
        res[name] = r["snaps_y"]
    rec["max_abs_diff"] = float((res["serial"] - res["scan"]).abs().max().item())
    rec["N"] = N
    out[f"z{zNo}"] = rec
print(json.dumps(out))
