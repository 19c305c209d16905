"""Update profiles/traffic.json from `ncu --set full` captures (needs only the ncu CLI, no GPU):

    python tools/ncu_traffic.py file.ncu-rep [file2.ncu-rep ...]

For every kernel found in the reports the mean dram__bytes_read.sum / dram__bytes_write.sum per launch is stored together with
the grid size and the report's name; bench.py reads the record at run time for `roofline.traffic`."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v*{"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


rec = json.load(open(OUT)) if os.path.exists(OUT) else {}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, ig, ib = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
                                               "launch__block_size"))
    by = {}
    for r in rows[2:]:
        by.setdefault(r[ik].split("(")[0], []).append((to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw]), r[ig], r[ib]))
    for k, v in by.items():
        rec[k] = {"dram_bytes_read": sum(x[0] for x in v)/len(v), "dram_bytes_write": sum(x[1] for x in v)/len(v), "launches": len(v),
                  "grid": v[0][2], "block": v[0][3], "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", ".txt")}
json.dump(rec, open(OUT, "w"), indent=1, sort_keys=True)
print(json.dumps(rec, indent=1, sort_keys=True))
