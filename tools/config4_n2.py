"""BASELINE config 4: dynamic model, zNo = 400, 1,048,576 reactors sharded over the GPUs of one box (131,072 per GPU at 8).

    python tools/config4_n2.py                                  # one rank = one GPU's shard of N_TOTAL/WORLD_DIV reactors
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_n2.py

Every rank solves its contiguous shard in chunks of CHUNK reactors (full snapshots of a chunk: CHUNK*5*7*400*8 B = 1.8 GB at
16,384; the outlet-node rows [x_i ; T] of every snapshot are kept, full profiles stay on the GPU that made them and are dropped
here), then ONE all-gather of the outlet rows.  A few reactors are checked against the oracle on rank 0.
Env: N_TOTAL (default 1048576), WORLD_DIV (pretend-world for single-GPU runs of one shard, default 8), ZNO, CHUNK, ORACLE=0/1.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyremot_b200 import cases, ensemble  # noqa: E402
from pyremot_b200.api import get_model  # noqa: E402

N_TOTAL = int(os.environ.get("N_TOTAL", 1048576))
ZNO, TNO, PERIOD = int(os.environ.get("ZNO", 400)), 5, 0.5
CHUNK = int(os.environ.get("CHUNK", 16384))
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
div = world if world > 1 else int(os.environ.get("WORLD_DIV", 8))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = ensemble.shard_bounds(N_TOTAL, rank, div)
mi = cases.readme_case("N2")
m = get_model(mi, device=local)
rows = cases.ensemble_rows(hi - lo, offset=lo)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()
n = m.n


def run(count):
    par = m.prolog(tab[:count])
    outlet = torch.empty((count, TNO*(n + 1)), dtype=torch.float64, device="cuda")
    status = torch.empty((count,), dtype=torch.int32, device="cuda")
    steps = torch.zeros((), dtype=torch.float64, device="cuda")
    for a in range(0, count, CHUNK):
        b = min(a + CHUNK, count)
        r = m.solve_n2(par[:, a:b].contiguous(), ZNO, TNO, PERIOD)
        data = m.unscale_n2(par[:, a:b].contiguous(), r["snaps_y"])                  # [chunk, tNo, n+1, zNo]
        outlet[a:b] = data[:, :, :, -1].reshape(b - a, -1)
        status[a:b] = r["status"]
        steps += r["stats"][0].double().sum()
        del r, data
    return outlet, status, steps


run(min(2*148, hi - lo))                                     # warm-up (module load, scratch allocation)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
outlet, status, steps = run(hi - lo)
full = ensemble.gather_rows(outlet, N_TOTAL, world, rank) if world > 1 else outlet
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
nbad = (status != 0).sum().to(torch.float64).reshape(1)
tot_steps = steps.reshape(1).clone()
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbad)
    dist.all_reduce(tot_steps)
if rank == 0:
    solved = N_TOTAL if world > 1 else hi - lo
    rec = {"config": "N2 zNo=%d tNo=%d period=%g" % (ZNO, TNO, PERIOD), "reactors_total": N_TOTAL, "gpus": world,
           "reactors_solved": solved, "reactors_per_gpu": hi - lo, "seconds": float(ms)*1e-3, "solves_per_s": solved/(float(ms)*1e-3),
           "failed": int(nbad), "mean_steps": float(tot_steps)/solved, "node_steps_per_s": float(tot_steps)*ZNO/(float(ms)*1e-3),
           "gathered_shape": list(full.shape), "finite": bool(torch.isfinite(full).all())}
    if os.environ.get("ORACLE", "1") != "0":
        from oracle import pbr_oracle as O
        idx = [0, (hi - lo)//2]
        ens = O.Ensemble(mi, {"T": rows["T"][idx], "P": rows["P"][idx], "conc": rows["conc"][idx], "Tm": rows["T"][idx]})
        errs = []
        _, data, _ = O.solve_n2(ens, ZNO, TNO, method="BDF", rtol=1e-8, atol=1e-11, sparse=True)      # [2, tNo, n+1, zNo] = [x ; T]
        for k, i in enumerate(idx):
            got = full[i].cpu().numpy().reshape(TNO, n + 1)
            errs.append(float(np.max(np.abs(got - data[k, :, :, -1])/np.abs(data[k, :, :, -1]))))
        rec["outlet_rel_err_vs_oracle"] = errs
    print(json.dumps(rec))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
