"""Multi-GPU host logic on CPU: the work list is sharded across ranks with no data-path collective
(para_gen.py --gpu style).  world_size-2 gloo group only checks coverage -- CPU only."""
import os

import pytest

from arap_flow_b200 import driver


def _items(n):
    return [tuple(f"/d/{k}_{i}.x" for k in ("rgb", "msk", "cstr", "flo", "wrgb", "wmsk")) for i in range(n)]


def test_list_file_round_trip(tmp_path):
    items = _items(5)
    p = str(tmp_path / "l.txt")
    driver.write_list_file(p, items)
    with open(p, "a") as f:
        f.write("too short line\n\n")
    assert driver.read_list_file(p) == items


@pytest.mark.parametrize("n,world", [(64, 1), (64, 2), (64, 4), (64, 8), (7, 4), (0, 2), (3, 8)])
def test_shards_partition_the_work(n, world):
    items = _items(n)
    shards = [driver.shard(items, r, world) for r in range(world)]
    assert sorted(sum(shards, [])) == sorted(items)
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        driver.shard(items, world, world)


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = driver.shard(_items(n), rank, world)
    # the only "communication" is the end-of-job barrier + bookkeeping of what each rank did
    ids = torch.zeros(n, dtype=torch.int64)
    for it in mine:
        ids[int(it[0].split("_")[1].split(".")[0])] = 1
    dist.all_reduce(ids)
    dist.barrier()
    q.put((rank, len(mine), ids.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_job_covers_every_unit_once():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, world, port = 13, 2, 29517
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[1] for r in res) == [6, 7]
    assert all(r[2] == [1] * n for r in res)  # every unit processed by exactly one rank
