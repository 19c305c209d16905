"""Result writers for ensembles: sharded ``.npz`` files, one per rank, replacing the list->txt/csv
helpers of `PyREMOT/library/saveResult.py:18-47` for batched results.

``save_shard`` writes what ``rmtExeBatch`` returns for one rank's slice together with the row range it covers;
``load_shards`` stitches a directory of shards back into full arrays in row order.  N2 results carry all ``tNo``
interval-end snapshots, so a saved N2 run is also the restart state of every reactor (the reference restarts
``solve_ivp`` from the previous interval's end state, `pbHomoReactor.py:3685`).
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

_KEYS = ("outlet", "dataYs", "snaps_y", "status", "stats")


def save_shard(directory, result, lo, hi, rank=0, meta=None):
    """Write rows [lo, hi) of an ensemble result. Returns the file path."""
    os.makedirs(directory, exist_ok=True)
    arrs = {k: np.asarray(result[k]) for k in _KEYS if result.get(k) is not None}
    for k, a in arrs.items():
        n = a.shape[1] if k == "stats" else a.shape[0]
        if n != hi - lo:
            raise ValueError(f"{k} has {n} rows, expected {hi - lo}")
    path = os.path.join(directory, f"shard_{rank:04d}.npz")
    np.savez_compressed(path, __rows__=np.array([lo, hi]), __meta__=np.array(json.dumps(meta or {})), **arrs)
    return path


def load_shards(directory):
    """Concatenate all shards of a directory in row order. Returns (result dict, meta of the first shard)."""
    files = sorted(glob.glob(os.path.join(directory, "shard_*.npz")))
    if not files:
        raise FileNotFoundError(f"no shards in {directory}")
    parts = []
    for f in files:
        z = np.load(f, allow_pickle=False)
        parts.append((int(z["__rows__"][0]), int(z["__rows__"][1]), z))
    parts.sort(key=lambda t: t[0])
    pos = 0
    for lo, hi, _ in parts:
        if lo != pos:
            raise ValueError(f"shards do not tile the row range: expected a shard starting at {pos}, found {lo}")
        pos = hi
    out = {}
    for k in _KEYS:
        if k in parts[0][2].files:
            out[k] = np.concatenate([p[2][k] for p in parts], axis=1 if k == "stats" else 0)
    return out, json.loads(str(parts[0][2]["__meta__"]))
