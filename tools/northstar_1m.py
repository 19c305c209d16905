"""North-star run: 1,048,576 steady-state methanol/DME reactors over the GPUs of one box (131,072 per GPU at 8 GPUs),
outlets gathered with one NCCL all-gather, a sample checked against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/northstar_1m.py
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
from pyremot_b200.api import get_model  # noqa: E402

N_TOTAL = int(os.environ.get("N_TOTAL", 1048576))
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = ensemble.shard_bounds(N_TOTAL, rank, world)
mi = cases.readme_case("N1")
m = get_model(mi, device=local)
rows = cases.ensemble_rows(hi - lo, offset=lo)
tab = torch.from_numpy(m.instance_table(mi, {"T": rows["T"], "P": rows["P"], "conc": rows["conc"], "Tm": rows["T"]})).cuda()


def run():
    par = m.prolog(tab)
    r = m.solve_n1(par, rtol=1e-9, atol=1e-13, profile=False)
    out = m.unscale_n1(par, r["outlet_y"][None])[:, :, 0].contiguous()
    return ensemble.gather_rows(out, N_TOTAL, world, rank), r


for _ in range(2):
    full, r = run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
full, r = run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
nbad = (r["status"] != 0).sum().to(torch.float32).reshape(1)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbad)
if rank == 0:
    from oracle import pbr_oracle as O
    idx = [0, 123457, 524288, N_TOTAL - 1]
    rr = cases.ensemble_rows(N_TOTAL)
    ens = O.Ensemble(mi, {"T": rr["T"][idx], "P": rr["P"][idx], "conc": rr["conc"][idx], "Tm": rr["T"][idx]})
    _, ref, _ = O.solve_n1(ens, method="BDF", rtol=1e-11, atol=1e-14)
    got = full[idx].cpu().numpy()
    err = float(np.max(np.abs(got - ref[:, :, -1])/np.abs(ref[:, :, -1])))
    print(json.dumps({"reactors": N_TOTAL, "gpus": world, "ms": float(ms), "solves_per_s": N_TOTAL/(float(ms)*1e-3),
                      "failed": int(nbad), "outlet_rel_err_vs_oracle_4_samples": err, "gathered_shape": list(full.shape)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
