"""Compare the saved results of tools/ab_round2.py legs: python tools/ab_compare.py base.npz other.npz [...]"""
import sys

import numpy as np

base = np.load(sys.argv[1])
for f in sys.argv[2:]:
    o = np.load(f)
    parts = []
    for k in base.files:
        if k in o.files:
            a, b = base[k], o[k]
            parts.append(f"{k} {np.max(np.abs(a - b)/(np.abs(a) + 1e-300 + 1e-12*np.max(np.abs(a)))):.2e}")
    print(f, " | ".join(parts))
