"""How much would ordering the reactors by a step-count predictor buy K2?  Times 65,536 solves in table order, ordered by inlet T,
by P, by the true step count of a previous run (upper bound), and block-interleaved longest-first.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N = 65536
mi = cases.readme_case("N1")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m.prolog(tab)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(p):
    ts = []
    for k in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m.solve_n1(p, rtol=1e-9, atol=1e-13, n_out=101, profile=False)
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


ms0, r = timed(par)
steps = r["stats"][0].cpu().numpy().astype(float)
nfev = r["stats"][2].cpu().numpy().astype(float)
out = {"table_order_ms": ms0, "steps_mean": steps.mean(), "steps_std": steps.std(), "steps_max": steps.max()}
feats = {"T": rows["T"], "P": rows["P"], "H2COx": rows["H2COx"], "CO2COx": rows["CO2COx"]}
out["corr_with_steps"] = {k: float(np.corrcoef(v, steps)[0, 1]) for k, v in feats.items()}
A = np.stack([np.ones(N)] + [v for v in feats.values()] + [v*w for i, v in enumerate(feats.values()) for w in list(feats.values())[i:]], axis=1)
coef, *_ = np.linalg.lstsq(A, steps, rcond=None)
pred = A @ coef
out["quadratic_fit_r2"] = float(1 - ((steps - pred)**2).sum()/((steps - steps.mean())**2).sum())
for name, key in (("by_T", rows["T"]), ("by_P", rows["P"]), ("by_true_steps", steps), ("by_quadratic_fit", pred)):
    order = torch.from_numpy(np.argsort(key, kind="stable")).cuda()
    ms, r2 = timed(par[:, order].contiguous())
    out[name + "_ms"] = ms
    # longest blocks first: reverse order so that the second wave gets the short ones
    order = torch.from_numpy(np.argsort(-key, kind="stable")).cuda()
    ms, r2 = timed(par[:, order].contiguous())
    out[name + "_desc_ms"] = ms
print(json.dumps(out))
