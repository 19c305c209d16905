"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group."""
import os
import subprocess
import sys

import numpy as np

from pyremot_b200 import ensemble

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    for N in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            b = [ensemble.shard_bounds(N, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i+1][0] for i in range(world-1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_deal_balances_and_inverts():
    rng = np.random.default_rng(0)
    proxy = rng.random(1001)
    for world in (2, 4, 8):
        perm = ensemble.deal(proxy, world)
        assert sorted(perm.tolist()) == list(range(1001))
        means = [proxy[perm[slice(*ensemble.shard_bounds(1001, r, world))]].mean() for r in range(world)]
        assert max(means) - min(means) < 0.01          # every rank gets the same mix
        rows = np.arange(1001)[:, None]*np.ones((1, 3))
        assert np.array_equal(ensemble.undeal(perm, rows[perm]), rows)


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from pyremot_b200 import ensemble
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, K = 11, 8                      # odd N: shards of 6 and 5 rows
table = np.arange(N*K, dtype=np.float64).reshape(N, K)
proxy = (np.arange(N)*7) % 11
perm = ensemble.deal(proxy, world)
lo, hi = ensemble.shard_bounds(N, rank, world)
local = torch.from_numpy(table[perm[lo:hi]]*2.0)          # stand-in for "solve my shard"
full = ensemble.gather_rows(local, N, world, rank)
out = ensemble.undeal(perm, full.numpy())
assert np.array_equal(out, table*2.0), (rank, out)
# fewer reactors than ranks: rank 1's shard is EMPTY (ADVICE r1: it must not crash or hang the gather)
N1 = 1
lo, hi = ensemble.shard_bounds(N1, rank, world)
assert (hi - lo) == (1 if rank == 0 else 0)
local = torch.from_numpy(table[lo:hi]*3.0)
full = ensemble.gather_rows(local, N1, world, rank)
assert full.shape == (1, K) and np.array_equal(full.numpy(), table[:1]*3.0), (rank, full)
print("rank", rank, "ok", flush=True)
dist.destroy_process_group()
'''


def test_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), ROOT],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_result_shards_roundtrip(tmp_path):
    from pyremot_b200 import io
    rng = np.random.default_rng(3)
    N = 11
    full = {"outlet": rng.random((N, 8)), "status": np.zeros(N, np.int32), "stats": rng.integers(0, 900, (4, N)).astype(np.int32)}
    for rank in range(2):
        lo, hi = ensemble.shard_bounds(N, rank, 2)
        part = {"outlet": full["outlet"][lo:hi], "status": full["status"][lo:hi], "stats": full["stats"][:, lo:hi]}
        io.save_shard(str(tmp_path), part, lo, hi, rank=rank, meta={"model": "N1", "rtol": 1e-9})
    got, meta = io.load_shards(str(tmp_path))
    assert meta == {"model": "N1", "rtol": 1e-9}
    for k in full:
        assert np.array_equal(got[k], full[k])
    os.remove(str(tmp_path / "shard_0000.npz"))
    try:
        io.load_shards(str(tmp_path))
        assert False, "a gap in the row range must be detected"
    except ValueError:
        pass


def test_chunk_bounds_tile_the_shard():
    for N in (0, 1, 5, 16384, 131072):
        for chunk in (1, 7, 16384):
            if N > 1000 and chunk < 100:
                continue
            b = ensemble.chunk_bounds(N, chunk)
            assert (not b and N == 0) or (b[0][0] == 0 and b[-1][1] == N)
            assert all(b[i][1] == b[i+1][0] for i in range(len(b) - 1)) and all(0 < hi - lo <= chunk for lo, hi in b)
