"""CPU, world_size 2, gloo: the N>1 host logic (scene sharding + the single all-gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mind_b200.distributed import shard_range, all_gather_predictions
    n_scenes, actors_per_scene = 7, [3, 1, 4, 1, 5, 9, 2]
    s, e = shard_range(n_scenes, rank, world)
    g = torch.Generator().manual_seed(0)
    cls_all = torch.rand(n_scenes, 6, generator=g)
    a_off = [0]
    for a in actors_per_scene:
        a_off.append(a_off[-1] + a)
    reg_all = torch.rand(a_off[-1], 6, 60, 5, generator=g)
    vel_all = torch.rand(a_off[-1], 6, 60, 2, generator=g)
    c, r, v = all_gather_predictions(cls_all[s:e], reg_all[a_off[s]:a_off[e]], vel_all[a_off[s]:a_off[e]])
    ok = torch.equal(c, cls_all) and torch.equal(r, reg_all) and torch.equal(v, vel_all)
    # equal shards take the single-collective fast path
    c2, _, _ = all_gather_predictions(cls_all[rank * 3:(rank + 1) * 3], reg_all[:2], vel_all[:2])
    ok = ok and torch.equal(c2, cls_all[:6])
    q.put((rank, ok, (s, e)))
    dist.destroy_process_group()


def test_shard_ranges():
    from mind_b200.distributed import shard_range
    for n in (0, 1, 7, 256, 2048):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1


def test_allgather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sorted(x[2] for x in res) == [(0, 4), (4, 7)]
