"""Device time of K2 (the fused N1 solver) alone over N reactors, for A/B runs of build knobs inside ONE gpurun call.

    [PBR_N1_REDUCED=0] [PBR_N1_BLOCK=224] [PBR_EXTRA_DEFINES=...] python tools/bench_k2.py [N] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases, engine  # noqa: E402

if os.environ.get("PBR_AB_LIB"):           # A/B against another build of libpbr.so inside the same gpurun call
    engine.LIB_PATH = os.path.abspath(os.environ["PBR_AB_LIB"])
from pyremot_b200.api import get_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mi = cases.readme_case("N1")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m.prolog(tab)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for k in range(reps + 2):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m.solve_n1(par, rtol=1e-9, atol=1e-13, n_out=101, profile=False)
    e1.record()
    torch.cuda.synchronize()
    if k >= 2:
        ts.append(e0.elapsed_time(e1))
st = r["stats"].cpu().numpy()
tag = " ".join(f"{k}={os.environ[k]}" for k in ("PBR_AB_LIB", "PBR_N1_REDUCED", "PBR_N1_BLOCK", "PBR_N1_NO_BALANCE", "PBR_EXTRA_DEFINES") if k in os.environ)
print(f"K2 N={N} [{tag}] ms min/med {min(ts):.3f}/{sorted(ts)[len(ts)//2]:.3f}  solves/s {N/min(ts)*1e3:.4g}  fails {int((r['status'] != 0).sum())} "
      f"steps/nfev/nlu {st[0].mean():.1f}/{st[2].mean():.1f}/{st[3].mean():.1f}  attr {m.kernel_attr('pbr_k2_solve_n1')}")
