"""N > 1 host-side logic on CPU: world_size-2 gloo run of the problem sharding + final controller gather
(the data path itself has no collective: shards are independent)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, PKG)
    from i2c_b200 import dist as idist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, du, dx = 5, 1, 2
    rng = np.random.default_rng(0)  # same full problem set on every rank
    K_full = torch.from_numpy(rng.normal(size=(B, T, du, dx)))
    k_full = torch.from_numpy(rng.normal(size=(B, T, du)))
    s_full = torch.from_numpy(rng.normal(size=(B, T, du, du)))
    cost_full = torch.from_numpy(rng.normal(size=(B,)))
    s, e = idist.shard_range(B, world, rank)
    K, k, sg, cost = idist.gather_controllers(K_full[s:e], k_full[s:e], s_full[s:e], B, extra=(cost_full[s:e],))
    ok = all(torch.equal(a, b) for a, b in ((K, K_full), (k, k_full), (sg, s_full), (cost, cost_full)))
    # gather onto rank 0 only (what the bench's final gather uses): shards land in their slices of the final tensor
    got = idist.gather_controllers(K_full[s:e], k_full[s:e], s_full[s:e], B, extra=(cost_full[s:e],), dst=0)
    if rank == 0:
        ok = ok and all(torch.equal(a, b) for a, b in zip(got, (K_full, k_full, s_full, cost_full)))
    else:
        ok = ok and all(g is None for g in got)
    torch.save({"ok": ok, "range": (s, e)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_shard_and_gather_world2(tmp_path, B):
    world, port = 2, 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert all(r["ok"] for r in res)
    assert res[0]["range"][0] == 0 and res[0]["range"][1] == res[1]["range"][0] and res[1]["range"][1] == B


def test_shard_range_properties():
    sys.path.insert(0, PKG)
    from i2c_b200 import dist as idist

    for B in (1, 5, 4096, 65537):
        for world in (1, 2, 3, 8):
            rs = [idist.shard_range(B, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [e - s for s, e in rs]
            assert max(sizes) - min(sizes) <= 1
