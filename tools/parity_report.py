"""Print the measured parity margins (GPU vs reference-generated golden vectors). Needs a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyremot_b200  # noqa: E402
from pyremot_b200 import cases, rmtExe, rmtExeBatch  # noqa: E402

G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
rel = lambda a, b, fl=0.0: float(np.max(np.abs(np.asarray(a, float) - b)/(np.abs(b) + fl)))

g = G("n1_solve")
for tag, pt in (("noniso", "non-iso-thermal"), ("iso", "iso-thermal")):
    d = rmtExe(cases.readme_case("N1", pt))["resModel"][0]
    print(f"N1 README {tag}: profile rel err vs converged reference {rel(d['dataYs'], g[f'{tag}_conv_dataYs']):.2e}; "
          f"vs reference at its default tolerance {rel(d['dataYs'], g[f'{tag}_default_dataYs']):.2e}")
gs = G("n1_sweep")
mi = cases.readme_case("N1")
out = rmtExeBatch(mi, {"T": gs["T"], "P": gs["P"], "conc": gs["conc"], "Tm": gs["T"]}, profile=True)
print(f"N1 sweep (32 rows x 101 points): max profile rel err {rel(out['dataYs'], gs['dataYs']):.2e}, outlet {rel(out['outlet'], gs['outlet']):.2e}")
for rt in (1e-8, 1e-10):
    o = rmtExeBatch(mi, {"T": gs["T"], "P": gs["P"], "conc": gs["conc"], "Tm": gs["T"]}, profile=True, rtol=rt, atol=rt*1e-4)
    print(f"   at rtol {rt:g}: profile {rel(o['dataYs'], gs['dataYs']):.2e}  mean steps {o['stats'][0].mean():.0f}")
g2 = G("n2_solve")
for tag, mk, zNo in (("readme_z20", lambda: cases.readme_case("N2"), 20), ("dme_z20", cases.dme_n2_case, 20),
                     ("ch4_z10", cases.ch4_n2_case, 10), ("readme_iso_z10", lambda: cases.readme_case("N2", "iso-thermal"), 10)):
    pyremot_b200.solverSetting["N2"]["zNo"] = zNo
    r = rmtExe(mk())["resModel"]["dataPack"]
    ys = np.array([d["dataYs"] for d in r])
    print(f"N2 {tag}: snapshots rel err vs tightened reference {rel(ys, g2[f'{tag}_conv_dataYs']):.2e}; vs default-tolerance reference {rel(ys, g2[f'{tag}_default_dataYs']):.2e}")
