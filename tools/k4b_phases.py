"""Where a step attempt of the block-per-reactor N2 solver spends its cycles: the measurement build PBR_K4B_PROFILE=1 records thread
0's clock at the phase boundaries (bookkeeping = predictor + error test + history update + order selection | Jacobian blocks +
inverses (pbr_k4b_eval<true>) | march | rest of the Newton loop = residual vector, update, norm reduction).  One wave of config 3
(592 reactors, zNo 100) and of the config-4 slice (148 reactors, zNo 400).

    python tools/k4b_phases.py > gpurun_out/r2_k4b_phases.json
"""
import json
import os
import sys

os.environ["PBR_EXTRA_DEFINES"] = ";".join(x for x in (os.environ.get("PBR_EXTRA_DEFINES", ""), "PBR_K4B_PROFILE=1") if x)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

mi = cases.readme_case("N2")
m = get_model(mi)
out = {}
_plain = []


def get_plain():
    """the same model without the measurement define (its `stats` are the step counters)"""
    if not _plain:
        from pyremot_b200 import engine
        saved = os.environ.pop("PBR_EXTRA_DEFINES")
        try:
            _plain.append(engine.Model(mi, device=0))
        finally:
            os.environ["PBR_EXTRA_DEFINES"] = saved
    return _plain[0]

for tag, N, zNo in (("config3_z100", 592, 100), ("config3_z100_one_block_per_sm", 148, 100), ("config4_z400", 148, 400)):
    rows = cases.ensemble_rows(N)
    tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
    par = m.prolog(tab)
    m.solve_n2(par, zNo, 5, 0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m.solve_n2(par, zNo, 5, 0.5)
    e1.record()
    torch.cuda.synchronize()
    st = r["stats"].cpu().numpy().astype(np.float64)*64.0
    tot = st.sum(axis=0)
    plain = get_plain().solve_n2(par, zNo, 5, 0.5)["stats"].cpu().numpy().astype(np.float64)     # step counters of the normal build
    out[tag] = {"ms": e0.elapsed_time(e1), "failed": int((r["status"] != 0).sum().item()),
                "cycles_per_reactor_mean": float(tot.mean()),
                "march_cycles_per_node": float(st[2].sum()/(plain[2].sum()*zNo)),
                "jacobian_cycles_per_attempt": float(st[1].sum()/plain[3].sum()),
                "share": {k: float(st[i].sum()/tot.sum()) for i, k in enumerate(("bookkeeping", "jacobian_blocks_and_inverses", "march", "newton_rest"))}}
print(json.dumps(out))
