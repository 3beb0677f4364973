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


def _fake_forward(net_in, geom):
    """stands in for the network: per-scene outputs that depend on every input of that scene only"""
    actors, a_idcs, lanes, l_idcs, rpe, tgt_nodes, tgt_rpe = net_in
    f = len(a_idcs)
    na, nl = actors.shape[0] // f, lanes.shape[0] // f
    scene = (actors.view(f, -1).sum(1) + lanes.view(f, -1).sum(1) + tgt_nodes.view(f, -1).sum(1) + tgt_rpe.view(f, -1).sum(1)
             + geom[0].view(f, -1).sum(1) - geom[1].view(f, -1).sum(1))
    cls = scene[:, None] * torch.arange(1, 7)[None]
    per_actor = actors.view(f * na, -1).sum(1) + scene.repeat_interleave(na)
    reg = per_actor[:, None, None, None] * torch.ones(1, 6, 60, 5)
    vel = per_actor[:, None, None, None] * torch.ones(1, 6, 60, 2) * 2
    return cls, reg, vel, None, None


def _fake_forward_packed(net_in, geom):
    """same outputs laid out like ScenePredNetB200.forward_packed: (cls, reg, vel, cov_vel, param, a_off, pack)"""
    from mind_b200.predictor import pack_layout
    cls, reg, vel, _, _ = _fake_forward(net_in, geom)
    B, A = cls.shape[0], reg.shape[0]
    (c0, cn), (r0, rn), (v0, vn) = pack_layout(B, A)
    pack = torch.zeros(v0 + vn)
    pack[c0:c0 + cn] = cls.reshape(-1); pack[r0:r0 + rn] = reg.reshape(-1); pack[v0:v0 + vn] = vel.reshape(-1)
    return (pack[c0:c0 + cn].view(B, 6), pack[r0:r0 + rn].view(A, 6, 60, 5), pack[v0:v0 + vn].view(A, 6, 60, 2), None, None, None, pack)


def _level(F, na, nl, seed):
    g = torch.Generator().manual_seed(seed)
    net_in = (torch.rand(F * na, 14, 48, generator=g), [range(i * na, (i + 1) * na) for i in range(F)],
              torch.rand(F * nl, 10, 16, generator=g), [range(i * nl, (i + 1) * nl) for i in range(F)], None,
              torch.rand(F, 10, 16, generator=g), torch.rand(F, 20, generator=g))
    geom = (torch.rand(F * (na + nl), 2, generator=g), torch.rand(F * (na + nl), 2, generator=g))
    return net_in, geom


def _tree_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mind_b200.distributed import sharded_level_forward
    ok = True
    for F, na, nl in ((1, 5, 7), (6, 5, 7), (7, 3, 2), (36, 2, 3)):          # 1 < world: replicated; 7: uneven shards
        net_in, geom = _level(F, na, nl, 100 + F)
        want = _fake_forward(net_in, geom)
        got = sharded_level_forward(_fake_forward, net_in, geom, F)
        ok = ok and all(torch.equal(a, b) for a, b in zip(got, want[:3]))
        # the predictor's output form: cls | reg | vel as views of ONE buffer -> equal shards take the single packed all-gather
        got = sharded_level_forward(_fake_forward_packed, net_in, geom, F)
        ok = ok and all(torch.equal(a, b) for a, b in zip(got, want[:3]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_level_shards_are_views_and_cover_the_frontier():
    from mind_b200.distributed import shard_level_inputs
    net_in, geom = _level(7, 3, 2, 1)
    seen = []
    for r in range(3):
        sub, g, (s, e) = shard_level_inputs(net_in, geom, 7, r, 3)
        seen += list(range(s, e))
        assert sub[0].data_ptr() == net_in[0][s * 3:].data_ptr() and sub[0].shape[0] == (e - s) * 3
        assert len(sub[1]) == e - s and len(sub[3]) == e - s and sub[5].shape[0] == e - s and g[0].shape[0] == (e - s) * 5
    assert seen == list(range(7))


def test_tree_level_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_tree_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
