"""Standalone RHS kernels: evals/s and achieved GB/s (K3 N2 RHS at zNo in argv, K_rhs_n1).  python tools/bench_rhs.py [zNo] [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

zNo = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mi2 = cases.readme_case("N2")
m2 = get_model(mi2)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m2.instance_table(mi2, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m2.prolog(tab)
Y = torch.empty((N, m2.nv2, zNo), dtype=torch.float64, device="cuda")
Y[:, :m2.n, :] = par[16:16 + m2.n].t()[:, :, None]
Y[:, m2.n, :] = torch.linspace(0, 0.2, zNo, device="cuda", dtype=torch.float64)[None, :]
dY = torch.empty_like(Y)
for _ in range(3):
    m2.rhs_n2(par, Y, out=dY)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    m2.rhs_n2(par, Y, out=dY)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/reps
gbs = 16.0*m2.nv2*zNo*N/(ms*1e-3)/1e9
out = {"kernel": "pbr_k3_rhs_n2", "zNo": zNo, "N": N, "ms": ms, "GBs": gbs, "node_evals_per_s": N*zNo/(ms*1e-3),
       "attr": m2.kernel_attr("pbr_k3_rhs_n2"), "defines": os.environ.get("PBR_EXTRA_DEFINES", "")}
# N1 RHS, thread per instance
N1 = 4*1024*1024
y = torch.rand((m2.nv1, N1), dtype=torch.float64, device="cuda")*0.1 + 0.5
y[m2.n] = 0.99
y[m2.n + 1] = 0.1
tab1 = tab[:1].repeat(N1, 1).contiguous()
par1 = m2.prolog(tab1)
f1 = torch.empty_like(y)
for _ in range(3):
    m2.rhs_n1(par1, y, out=f1)
e0.record()
for _ in range(reps):
    m2.rhs_n1(par1, y, out=f1)
e1.record()
torch.cuda.synchronize()
ms1 = e0.elapsed_time(e1)/reps
out["rhs_n1"] = {"N": N1, "ms": ms1, "evals_per_s": N1/(ms1*1e-3), "GBs_state": 16.0*m2.nv1*N1/(ms1*1e-3)/1e9,
                 "GBs_with_par": (16.0*m2.nv1 + 8.0*m2.k_par)*N1/(ms1*1e-3)/1e9}
print(json.dumps(out))
