"""BASELINE.json configs[4]: mixed Quadrotor (hovering_control) + MetaMaze2D random-action rollout, envs sharded over
the GPUs of one node, one NCCL all-gather of every T-step trajectory chunk for the learner.

    python scripts/bench_mixed.py                                  # 1 GPU: 16 384 + 16 384 envs
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/bench_mixed.py                                     # 8 GPUs: 262 144 envs

Prints one JSON line (rank 0): env-steps/s without and with the gather, and the gather's bus bandwidth.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from metagym_b200 import BatchedMetaMaze2D, BatchedQuadrotor, MazeTaskSampler
from metagym_b200.rollout import MulticastArena, PeerArena, RolloutArena, all_gather_rollout, rollout_bytes, shard_range

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_QUAD = N_MAZE = 16384          # per GPU (131 072 + 131 072 over 8 GPUs)
T, CHUNKS = 32, 8
FUSED_MAZE = os.environ.get("MIXED_FUSED_MAZE", "1") != "0"   # mgb_maze_rollout vs T single mgb_maze_step calls
qbase, _ = shard_range(N_QUAD * world, rank, world)
mbase, _ = shard_range(N_MAZE * world, rank, world)
quad = BatchedQuadrotor(task="hovering_control", dt=0.01, nt=1000, num_envs=N_QUAD, device=local, squeeze=False,
                        auto_reset=True, env_index_base=qbase)
rs = np.random.RandomState(0)
tasks = [MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs) for _ in range(64)]
maze = BatchedMetaMaze2D(max_steps=200, task_type="ESCAPE", view_grid=1, num_envs=N_MAZE, device=local, squeeze=False,
                         auto_reset=True, env_index_base=mbase)
maze.set_task(tasks)
quad.reset()
maze.reset()
g = torch.Generator(device=dev).manual_seed(rank)
FIELDS = {
    "q_obs": ((T, N_QUAD, 16), torch.float32), "q_act": ((T, N_QUAD, 4), torch.float32),
    "q_rew": ((T, N_QUAD), torch.float32), "q_done": ((T, N_QUAD), torch.uint8),
    "m_obs": ((T, N_MAZE, 3, 3), torch.float32), "m_act": ((T, N_MAZE), torch.int32),
    "m_rew": ((T, N_MAZE), torch.float64), "m_done": ((T, N_MAZE), torch.uint8),
}
arenas = [RolloutArena(FIELDS, dev), RolloutArena(FIELDS, dev)]     # double buffer: collect k+1 while k is gathered
chunk = arenas[0].views


def collect(chunk=chunk):
    # quadrotor: T fused steps, device-drawn U(0.1, 15) actions; maze: T fused steps, device-drawn uniform {0..3}
    # actions (MIXED_FUSED_MAZE=0: T single steps with torch-drawn actions)
    quad.rollout(T, actions=None, act_seed=7, out={"obs": chunk["q_obs"], "rew": chunk["q_rew"], "done": chunk["q_done"],
                                                   "act": chunk["q_act"]})
    if FUSED_MAZE:
        maze.rollout(T, actions=None, act_seed=9, out={"obs": chunk["m_obs"], "rew": chunk["m_rew"],
                                                       "done": chunk["m_done"], "act": chunk["m_act"]})
        return
    chunk["m_act"].copy_(torch.randint(0, 4, (T, N_MAZE), device=dev, generator=g, dtype=torch.int32))
    for t in range(T):
        o, r, d, _ = maze.step(chunk["m_act"][t])
        chunk["m_obs"][t].copy_(o); chunk["m_rew"][t].copy_(r); chunk["m_done"][t].copy_(d)


def timed(fn, reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / reps


collect()
ms_collect = timed(collect, CHUNKS)
gathered, scratch = {}, {}


def collect_and_gather():
    collect()
    gathered.update(all_gather_rollout(chunk, scratch=scratch))


for _ in range(3):
    collect_and_gather()
ms_both = timed(collect_and_gather, CHUNKS)


def collect_and_gather_arena():
    collect(arenas[0].views)
    gathered.update(arenas[0].all_gather()[0])


pending = [None]
tick = [0]


def pipelined():
    ar = arenas[tick[0] & 1]
    tick[0] += 1
    collect(ar.views)                     # overlaps the previous chunk's gather (other arena, NCCL's stream)
    if pending[0] is not None:
        pending[0].wait()
    views, pending[0] = ar.all_gather(async_op=True)
    gathered.update(views)


for _ in range(3):
    collect_and_gather_arena()
ms_arena = timed(collect_and_gather_arena, CHUNKS)
ms_arena_gather = timed(lambda: arenas[0].all_gather(), CHUNKS)
for _ in range(4):
    pipelined()
ms_pipe = timed(pipelined, CHUNKS * 2)
if pending[0] is not None:
    pending[0].wait()
# ---- kernel-side gather: the rollout kernels store every output into all ranks' receive arenas (NVLink peer stores)
peers = [PeerArena(FIELDS, dev), PeerArena(FIELDS, dev)]
ptick = [0]


def collect_peer():
    ar = peers[ptick[0] & 1]
    ptick[0] += 1
    ar.attach(quad, maze)
    collect(ar.views)
    gathered.update(ar.sync())
    return ar


ar = collect_peer()
torch.cuda.synchronize(dev)
peer_ok = True
if world > 1:
    check = torch.empty(world * ar.nbytes, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(check, ar.buf.contiguous())
    peer_ok = bool(torch.equal(check, ar._recv))          # peer-written arena == NCCL all-gather of the same chunk
    assert peer_ok, "peer-written arena differs from the NCCL all-gather"
for _ in range(3):
    collect_peer()
ms_peer = timed(collect_peer, CHUNKS * 2)
# four arenas in rotation, rendezvous asynchronous: chunk k+1 rolls out while the rendezvous of chunk k completes.
# Stream order per rank: rollout(k), [rendezvous(k) on NCCL's stream], wait rendezvous(k-1), consume(k-1), rollout(k+1)...
# rendezvous(k-1) complete  =>  every rank has consumed chunk k-3  =>  rollout(k+1) may overwrite the arena of chunk k-3.
while len(peers) < 4:
    peers.append(PeerArena(FIELDS, dev))
works = [None] * 4


def collect_peer_async():
    i = ptick[0] % 4
    ptick[0] += 1
    ar = peers[i]
    ar.attach(quad, maze)
    collect(ar.views)
    _, works[i] = ar.sync(async_op=True)
    j = (i - 1) % 4
    if works[j] is not None:
        works[j].wait()
        works[j] = None
        gathered.update(peers[j].gathered)      # the learner would read chunk k-1 here


for _ in range(6):
    collect_peer_async()
ms_peer_async = timed(collect_peer_async, CHUNKS * 3)
for w_ in works:
    if w_ is not None:
        w_.wait()
torch.cuda.synchronize(dev)
quad.set_mirrors([])
maze.set_mirrors([])
# ---- NVSwitch multicast: each output stored once with multimem.st, replicated by the switch into every rank's arena
ms_mc, mc_ok, mc_err = None, None, None
if world > 1 and os.environ.get("MIXED_MULTICAST", "1") != "0":
    try:
        mcs = [MulticastArena(FIELDS, dev) for _ in range(4)]
    except Exception as e:                       # no NVLS on this box: report, do not fail the other measurements
        mcs, mc_err = None, repr(e)[:200]
    if mcs:
        mtick = [0]
        mworks = [None] * 4

        def collect_mc():
            i = mtick[0] % 4
            mtick[0] += 1
            ar = mcs[i]
            ar.attach(quad, maze)
            collect(ar.views)
            _, mworks[i] = ar.sync(async_op=True)
            j = (i - 1) % 4
            if mworks[j] is not None:
                mworks[j].wait()
                mworks[j] = None
                gathered.update(mcs[j].gathered)

        collect_mc()
        mworks[0].wait()
        mworks[0] = None
        torch.cuda.synchronize(dev)
        check = torch.empty(world * mcs[0].nbytes, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(check, mcs[0].buf.contiguous())
        mc_ok = bool(torch.equal(check, mcs[0]._recv)) and float(mcs[0].gathered["q_obs"].abs().sum()) > 0
        assert mc_ok, "multicast-written arena differs from the NCCL all-gather"
        for _ in range(7):
            collect_mc()
        ms_mc = timed(collect_mc, CHUNKS * 3)
        for w_ in mworks:
            if w_ is not None:
                w_.wait()
        torch.cuda.synchronize(dev)
        quad.set_multicast(0)
        maze.set_multicast(0)
ms_gather = timed(lambda: all_gather_rollout(chunk, scratch=scratch), CHUNKS)
steps = (N_QUAD + N_MAZE) * world * T
nbytes = rollout_bytes(chunk)
if rank == 0:
    print(json.dumps({
        "config": "mixed quadrotor hovering_control + MetaMaze2D ESCAPE, %d envs over %d GPU(s), T=%d chunks" %
                  ((N_QUAD + N_MAZE) * world, world, T),
        "env_steps_per_s_no_gather": steps / (ms_collect * 1e-3),
        "env_steps_per_s_with_gather": steps / (ms_both * 1e-3),
        "env_steps_per_s_with_arena_gather": steps / (ms_arena * 1e-3),
        "env_steps_per_s_pipelined_arena_gather": steps / (ms_pipe * 1e-3),
        "env_steps_per_s_kernel_side_gather": steps / (ms_peer * 1e-3), "kernel_side_gather_equals_nccl": peer_ok,
        "env_steps_per_s_kernel_side_gather_async_rendezvous": steps / (ms_peer_async * 1e-3),
        "env_steps_per_s_multicast_gather": steps / (ms_mc * 1e-3) if ms_mc else None,
        "multicast_gather_equals_nccl": mc_ok, "multicast_error": mc_err,
        "arena_allgather_ms": ms_arena_gather,
        "arena_allgather_busbw_GBps": arenas[0].nbytes * (world - 1) / (ms_arena_gather * 1e-3) / 1e9 if world > 1 else None,
        "allgather_ms": ms_gather, "chunk_bytes_per_rank": nbytes,
        "allgather_busbw_GBps": nbytes * (world - 1) / (ms_gather * 1e-3) / 1e9 if world > 1 else None,
        "fused_maze_rollout": FUSED_MAZE, "gathered_envs": int(gathered["q_obs"].shape[0] * gathered["q_obs"].shape[2]) + int(gathered["m_obs"].shape[0] * gathered["m_obs"].shape[2]), "n_gpus": world}), flush=True)
for a_ in peers:
    a_.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
