"""Does an L2 flush before the launch change K2's duration? (bench.py flushes between timed iterations.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyremot_b200 import cases  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N = 65536
mi = cases.readme_case("N1")
m = get_model(mi)
rows = cases.ensemble_rows(N)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
par = m.prolog(tab)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
for mode in ("none", "flush", "flush+sleep", "prolog-before", "none"):
    for _ in range(2):
        r = m.solve_n1(par, profile=False)
    torch.cuda.synchronize()
    ts = []
    for k in range(5):
        if mode.startswith("flush"):
            flush.fill_(k)
        if mode == "flush+sleep":
            torch.cuda.synchronize()
            torch.cuda._sleep(20_000_000)
        if mode == "prolog-before":
            par = m.prolog(tab)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m.solve_n1(par, profile=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{mode:14s} K2 ms: " + " ".join(f"{t:.2f}" for t in ts))
