"""BASELINE config 4: dynamic model, zNo = 400, 1,048,576 reactors sharded over the GPUs of one box (131,072 per GPU at 8).

    python tools/config4_n2.py                                  # one GPU: the first shard (N_TOTAL/WORLD_DIV reactors)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_n2.py

Runs `pyremot_b200.ensemble.solve_n2_sharded` (chunked per-rank solve, outlet-node rows kept, ONE all-gather) and checks two
reactors against the oracle on rank 0.  Env: N_TOTAL (default 1048576), WORLD_DIV (single-GPU runs solve 1/WORLD_DIV of the
table, default 8), ZNO (400), CHUNK (16384), ORACLE=0/1.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyremot_b200 import cases, ensemble  # noqa: E402

N_TOTAL = int(os.environ.get("N_TOTAL", 1048576))
ZNO, TNO = int(os.environ.get("ZNO", 400)), 5
CHUNK = int(os.environ.get("CHUNK", 16384))
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = N_TOTAL if world > 1 else N_TOTAL//int(os.environ.get("WORLD_DIV", 8))
mi = cases.readme_case("N2")                     # period 0.5 s, start-up from the feed-filled bed
rows = cases.ensemble_rows(N)
sweep = {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]}

ensemble.solve_n2_sharded(mi, {k: v[:2*148*world] for k, v in sweep.items()}, ZNO, TNO, chunk=CHUNK, device=local)   # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
full, status, mean_steps = ensemble.solve_n2_sharded(mi, sweep, ZNO, TNO, chunk=CHUNK, device=local)
torch.cuda.synchronize()
sec = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
nbad = torch.tensor([float((status != 0).sum())], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbad)
if rank == 0:
    rec = {"config": "N2 zNo=%d tNo=%d period=%g" % (ZNO, TNO, mi["operating-conditions"]["period"]), "gpus": world,
           "reactors": N, "reactors_per_gpu": N//world, "seconds_incl_host_table_and_gather": float(sec),
           "solves_per_s": N/float(sec), "failed": int(nbad), "mean_steps_rank0": mean_steps,
           "node_steps_per_s_per_gpu": mean_steps*ZNO*(N//world)/float(sec),
           "gathered_shape": list(full.shape), "finite": bool(np.isfinite(full).all())}
    if os.environ.get("ORACLE", "1") != "0":
        from oracle import pbr_oracle as O
        idx = [0, N//world//2]
        ens = O.Ensemble(mi, {k: v[idx] for k, v in sweep.items()})
        _, data, _ = O.solve_n2(ens, ZNO, TNO, method="BDF", rtol=1e-8, atol=1e-11, sparse=True)      # [2, tNo, n+1, zNo] = [x ; T]
        rec["outlet_rel_err_vs_oracle"] = [float(np.max(np.abs(full[i] - data[k, :, :, -1])/np.abs(data[k, :, :, -1])))
                                           for k, i in enumerate(idx)]
    print(json.dumps(rec))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
