"""N2 ensemble solve timing (BASELINE config 3 by default): python tools/bench_n2.py [N] [zNo] [tNo] [period] [rtol]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
zNo = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tNo = int(sys.argv[3]) if len(sys.argv) > 3 else 5
period = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
rtol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-9
mi = cases.readme_case("N2")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m.prolog(tab)
r = m.solve_n2(par, zNo, tNo, period, rtol=rtol, atol=rtol*1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 2
e0.record()
for _ in range(reps):
    r = m.solve_n2(par, zNo, tNo, period, rtol=rtol, atol=rtol*1e-4)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/reps
st = r["stats"].cpu().numpy()
if os.environ.get("PBR_BENCH_SAVE"):      # first 64 reactors' snapshots, for A/B comparisons between kernel variants
    import numpy as np
    np.save(os.environ["PBR_BENCH_SAVE"], r["snaps_y"][:64].cpu().numpy())
kname = "pbr_k4b_solve_n2" if zNo <= 56 else "pbr_k4bs128_solve_n2" if zNo <= 128 else "pbr_k4bs_solve_n2" if zNo <= 256 else "pbr_k4d_solve_n2 / pbr_k4c_solve_n2" if zNo <= 512 else "pbr_k4_solve_n2"
print(json.dumps({"kernel": kname + " (launcher's default for this grid)", "N": N, "zNo": zNo, "tNo": tNo, "period": period, "rtol": rtol, "ms": ms,
                  "solves_per_s": N/(ms*1e-3), "failed": int((r["status"] != 0).sum()), "mean_steps": float(st[0].mean()),
                  "mean_nfev": float(st[2].mean()), "mean_nlu": float(st[3].mean()),
                  "node_steps_per_s": float(st[0].sum())*zNo/(ms*1e-3), "attr": m.kernel_attr(kname.split(" ")[0])}))
