"""A/B of compile-time kernel variants (PBR_EXTRA_DEFINES) on the N2 ensemble solve, in one process group on one GPU.

  python tools/ab_variants.py --precompile            # here (no GPU): NVRTC-compile every variant into pyremot_b200/_cache
  python tools/ab_variants.py [N] [zNo]               # on the GPU box: time each variant, compare snapshots with the default
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "default": "",
    "restart_every_interval": "PBR_N2_RESTART=1",
    "k4b_scan_march": "PBR_K4B_MARCH_SCAN=1",
    "k4c_serial_march": "PBR_K4C_MARCH_SCAN=0",
}
# measured and recorded in profiles/r1l_k4*_variants_*.jsonl: PBR_K4B_MARCH_ROWSPLIT=0, PBR_K4B_MARCH_SCAN=1,
# PBR_K4B_INVERT_REG=1, PBR_K4B_FLAG_REDUCTION=0 and their combinations; in profiles/r1m_k4*_variants_*.jsonl:
# PBR_K4B_MARCH_V2=1 (adopted), PBR_K4B_JAC_REUSE=1 with PBR_K4B_JAC_CTOL=0 / 0.2 (rejected: 1.5x slower);
# in profiles/r1n_*: PBR_K4B_EVAL_OUTLINE=1 (rejected), K4d on chip (adopted).  The code of the rejected variants has
# been removed from pbr_device.cuh; the switches that remain are the ones listed above


def env_for(defs):
    e = dict(os.environ)
    if defs:
        e["PBR_EXTRA_DEFINES"] = defs
    else:
        e.pop("PBR_EXTRA_DEFINES", None)
    return e


if "--precompile" in sys.argv:
    code = ("import sys; sys.path.insert(0, %r)\nfrom pyremot_b200 import cases, engine\n"
            "m = engine.Model(cases.readme_case('N2'), device=None); m.close()\n" % ROOT)
    procs = [(k, subprocess.Popen([sys.executable, "-c", code], env=env_for(d))) for k, d in VARIANTS.items()]
    for k, p in procs:
        print(k, "rc", p.wait())
    sys.exit(0)

import numpy as np  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = args[0] if args else "4096"
zNo = args[1] if len(args) > 1 else "100"
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
ref = None
for k, d in VARIANTS.items():
    e = env_for(d)
    e["PBR_BENCH_SAVE"] = os.path.join(out, "ab_%s.npy" % k.replace("+", "_"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_n2.py"), N, zNo], env=e, capture_output=True, text=True)
    if r.returncode != 0:
        print(json.dumps({"variant": k, "error": r.stderr[-400:]}))
        continue
    j = json.loads(r.stdout.strip().splitlines()[-1])
    s = np.load(e["PBR_BENCH_SAVE"])
    os.remove(e["PBR_BENCH_SAVE"])
    if ref is None:
        ref = s
    dev = float(np.max(np.abs(s - ref)/(1e-13 + 1e-9*np.abs(ref))))      # in tolerance units of the solve
    print(json.dumps({"variant": k, "ms": round(j["ms"], 2), "solves_per_s": round(j["solves_per_s"]), "failed": j["failed"],
                      "mean_steps": round(j["mean_steps"], 1), "mean_nfev": round(j["mean_nfev"], 1),
                      "dev_vs_default_tol_units": round(dev, 2), "regs": j["attr"]}), flush=True)
