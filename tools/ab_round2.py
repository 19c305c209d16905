"""One A/B leg (build knobs come from the environment: PBR_EXTRA_DEFINES, PBR_RATEC_OPT, PBR_N1_BLOCK, ...): device times of
K3 (N2 RHS, zNo 100 / 400), pbr_k_rhs_n1, K2 (65,536 N1 solves), K4b (config 3) and K4d (config-4 slice), one JSON line;
a few results are saved to gpurun_out/ab/<tag>.npz so that the legs can be compared with each other afterwards.

    PBR_AB_TAG=name [knobs] python tools/ab_round2.py [what]      # what: comma list of k3,rhs1,k2,k4b,k4d (default all)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

what = (sys.argv[1] if len(sys.argv) > 1 else "k3,rhs1,k2,k4b,k4d").split(",")
tag = os.environ.get("PBR_AB_TAG", "default")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"tag": tag, "env": {k: os.environ[k] for k in os.environ if k.startswith("PBR_") and k != "PBR_AB_TAG"}}
save = {}


def timed(fn, reps=5, warm=2):
    ts = []
    for k in range(reps + warm):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


def table(m, mi, N, offset=0):
    rows = cases.ensemble_rows(N, offset=offset)
    return torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()


mi1, mi2 = cases.readme_case("N1"), cases.readme_case("N2")
if any(w in what for w in ("k3", "k4b", "k4d")):
    m2 = get_model(mi2)
if "k3" in what:
    for zNo, N in ((100, 65536), (400, 16384)):
        par = m2.prolog(table(m2, mi2, N))
        Y = torch.empty((N, m2.nv2, zNo), dtype=torch.float64, device="cuda")
        Y[:, :m2.n, :] = par[16:16 + m2.n].t()[:, :, None]*(1.0 - 0.3*torch.linspace(0, 1, zNo, device="cuda", dtype=torch.float64))[None, None, :]
        Y[:, m2.n, :] = torch.linspace(0, 0.2, zNo, device="cuda", dtype=torch.float64)[None, :]
        dY = torch.empty_like(Y)
        ms, _ = timed(lambda: m2.rhs_n2(par, Y, out=dY), reps=10, warm=3)
        out[f"k3_z{zNo}"] = {"ms": ms, "GBs": 16.0*m2.nv2*zNo*N/(ms*1e-3)/1e9, "node_evals_per_s": N*zNo/(ms*1e-3), "attr": m2.kernel_attr("pbr_k3_rhs_n2")}
        save[f"k3_z{zNo}"] = dY[:32].cpu().numpy()
if "rhs1" in what:
    m1 = get_model(mi1)
    N1 = 4*1024*1024
    y = torch.rand((m1.nv1, N1), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))*0.1 + 0.5
    y[m1.n] = 0.99
    y[m1.n + 1] = 0.1
    par1 = m1.prolog(table(m1, mi1, 1).repeat(N1, 1).contiguous())
    f1 = torch.empty_like(y)
    ms, _ = timed(lambda: m1.rhs_n1(par1, y, out=f1), reps=10, warm=3)
    out["rhs_n1"] = {"ms": ms, "GBs": 8.0*(2*m1.nv1 + m1.k_par - m1.n - 1)*N1/(ms*1e-3)/1e9, "attr": m1.kernel_attr("pbr_k_rhs_n1")}
    save["rhs_n1"] = f1[:, :256].cpu().numpy()
if "k2" in what:
    m1 = get_model(mi1)
    N = int(os.environ.get("PBR_AB_N1", "65536"))
    par = m1.prolog(table(m1, mi1, N))
    ms, r = timed(lambda: m1.solve_n1(par, rtol=1e-9, atol=1e-13, n_out=101, profile=False))
    st = r["stats"].cpu().numpy()
    out["k2"] = {"ms": ms, "solves_per_s": N/(ms*1e-3), "failed": int((r["status"] != 0).sum()), "steps": float(st[0].mean()),
                 "nfev": float(st[2].mean()), "nlu": float(st[3].mean()), "attr": m1.kernel_attr("pbr_k2_solve_n1"), "block": m1.n1_block}
    save["k2_outlet"] = m1.unscale_n1(par, r["outlet_y"][None])[:512, :, 0].cpu().numpy()
    msp, rp = timed(lambda: m1.solve_n1(par[:, :8192].contiguous(), rtol=1e-9, atol=1e-13, n_out=101, profile=True), reps=2, warm=1)
    save["k2_prof"] = m1.unscale_n1(par[:, :8192].contiguous(), rp["prof_y"])[:64].cpu().numpy()
for key, N, zNo, kern in (("k4b", 4096, 100, "pbr_k4b_solve_n2"), ("k4d", 1184, 400, "pbr_k4d_solve_n2")):
    if key in what:
        par = m2.prolog(table(m2, mi2, N))
        ms, r = timed(lambda: m2.solve_n2(par, zNo, 5, 0.5, rtol=1e-9, atol=1e-13), reps=3, warm=1)
        st = r["stats"].cpu().numpy()
        out[key] = {"ms": ms, "solves_per_s": N/(ms*1e-3), "failed": int((r["status"] != 0).sum()), "steps": float(st[0].mean()),
                    "nfev": float(st[2].mean()), "attr": m2.kernel_attr(kern)}
        save[key + "_snaps"] = r["snaps_y"][:8].cpu().numpy()
os.makedirs("gpurun_out/ab", exist_ok=True)
np.savez_compressed(f"gpurun_out/ab/{tag}.npz", **save)
print(json.dumps(out), flush=True)
