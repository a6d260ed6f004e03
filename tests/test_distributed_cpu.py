"""CPU, world_size 2, gloo: the sharding + rollout all-gather host logic of metagym_b200.rollout."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metagym_b200.rollout import RolloutArena, all_gather_rollout, shard_range


def test_shard_range_partitions_everything():
    for n, w in [(65536, 8), (10, 3), (7, 8), (262144, 8), (1, 1)]:
        seen = []
        for r in range(w):
            base, cnt = shard_range(n, r, w)
            seen += list(range(base, base + cnt))
        assert seen == list(range(n))


def test_rollout_arena_layout_single_process():
    """Fields are carved out of one allocation at 256-byte offsets; without a process group the gather is the identity."""
    fields = {"obs": ((3, 5, 19), torch.float32), "done": ((3, 5), torch.uint8), "rew": ((3, 5), torch.float64),
              "act": ((3, 5), torch.int32)}
    ar = RolloutArena(fields, "cpu")
    base = ar.buf.data_ptr()
    offs = [ar[k].data_ptr() - base for k in fields]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert ar.payload_bytes() == 3 * 5 * (19 * 4 + 1 + 8 + 4) and ar.nbytes >= ar.payload_bytes()
    for k, (shape, dt) in fields.items():
        assert tuple(ar[k].shape) == shape and ar[k].dtype == dt and ar[k].is_contiguous()
    ar["obs"].fill_(2.5); ar["done"].fill_(1)
    views, work = ar.all_gather()
    assert work is None and tuple(views["obs"].shape) == (1, 3, 5, 19)
    assert views["obs"].data_ptr() == ar["obs"].data_ptr() and float(views["obs"].sum()) == 2.5 * 3 * 5 * 19
    flat = RolloutArena.ordered(views)
    assert tuple(flat["done"].shape) == (3, 5) and int(flat["done"].sum()) == 15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, n_local, D = 3, 5, 4
    base, cnt = shard_range(n_local * world, rank, world)
    assert cnt == n_local
    env_ids = torch.arange(base, base + cnt)
    # a synthetic "rollout" whose content encodes (t, global env id): gather must be ordered by global env index
    obs = (torch.arange(T).view(T, 1, 1) * 1000 + env_ids.view(1, -1, 1) * 10 + torch.arange(D).view(1, 1, D)).float()
    chunk = {"obs": obs, "rew": obs[..., 0].double(), "done": (obs[..., 0] % 20 == 0), "act": None}
    g = all_gather_rollout(chunk)
    ids = torch.arange(n_local * world)
    want = (torch.arange(T).view(T, 1, 1) * 1000 + ids.view(1, -1, 1) * 10 + torch.arange(D).view(1, 1, D)).float()
    ok = torch.equal(g["obs"], want) and torch.equal(g["rew"], want[..., 0].double())
    ok = ok and torch.equal(g["done"], want[..., 0] % 20 == 0) and g["done"].dtype == torch.bool and "act" not in g
    # the same chunk through the single-allocation arena: one collective, zero-copy [world, T, n, ...] views
    arena = RolloutArena({"obs": ((T, n_local, D), torch.float32), "rew": ((T, n_local), torch.float64),
                          "done": ((T, n_local), torch.uint8), "act": ((T, n_local), torch.int32)}, "cpu")
    arena["obs"].copy_(obs); arena["rew"].copy_(obs[..., 0].double()); arena["done"].copy_(chunk["done"])
    arena["act"].copy_(env_ids.view(1, -1).expand(T, -1))
    for async_op in (False, True):
        views, work = arena.all_gather(async_op=async_op)
        if work is not None:
            work.wait()
        ok = ok and tuple(views["obs"].shape) == (world, T, n_local, D)
        ok = ok and views["obs"].untyped_storage().data_ptr() == views["act"].untyped_storage().data_ptr()
        flat = RolloutArena.ordered(views)
        ok = ok and torch.equal(flat["obs"], want) and torch.equal(flat["rew"], want[..., 0].double())
        ok = ok and torch.equal(flat["done"].bool(), want[..., 0] % 20 == 0)
        ok = ok and torch.equal(flat["act"], ids.view(1, -1).expand(T, -1).int())
    ok = ok and arena.nbytes % 256 == 0 and arena.payload_bytes() == T * n_local * (D * 4 + 8 + 1 + 4)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_rollout_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
