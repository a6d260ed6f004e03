"""GPU tests of the kernel-side trajectory all-gather: output mirrors of the fused rollout kernels and the peer-memory
(cudaIpc) entry points of the C ABI.  One GPU is enough: a mirror is just `pointer + delta`, here a second arena on the
same device; the cross-process mapping is exercised with a child process on the same GPU.  The 2-GPU run over NVLink
is scripts/bench_mixed.py (it asserts the peer-written arenas equal an NCCL all-gather of the same chunk)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_mod(cuda_device):
    import torch
    return torch


def _arenas(torch, fields, k):
    from metagym_b200.rollout import RolloutArena
    return [RolloutArena(fields, "cuda:0") for _ in range(k)]


@pytest.mark.parametrize("n,task", [(1000, "velocity_control"), (77, "hovering_control")])
def test_quad_rollout_mirrors_equal_primary(torch_mod, n, task):
    torch = torch_mod
    from metagym_b200 import BatchedQuadrotor
    T = 12
    D = 19 if task == "velocity_control" else 16
    fields = {"obs": ((T, n, D), torch.float32), "act": ((T, n, 4), torch.float32), "rew": ((T, n), torch.float32),
              "done": ((T, n), torch.uint8)}
    a0, a1, a2, ref = _arenas(torch, fields, 4)

    def run(arena, mirrors):
        env = BatchedQuadrotor(task=task, num_envs=n, device=0, squeeze=False, auto_reset=True, nt=40, dt=0.01)
        env.reset()
        env.set_mirrors(mirrors)
        for _ in range(2):                       # two chunks: the second overwrites, the state hand-over is the usual one
            env.rollout(T, act_seed=5, out={"obs": arena["obs"], "rew": arena["rew"], "done": arena["done"],
                                            "act": arena["act"]})
        torch.cuda.synchronize()
        env.close()

    run(ref, [])
    run(a0, [a1.buf.data_ptr() - a0.buf.data_ptr(), a2.buf.data_ptr() - a0.buf.data_ptr()])
    assert torch.equal(a0.buf, ref.buf)          # mirroring does not change the primary outputs (bit-exact)
    assert torch.equal(a1.buf, a0.buf) and torch.equal(a2.buf, a0.buf)
    assert float(ref["obs"].abs().sum()) > 0


@pytest.mark.parametrize("n,view_grid,task_type", [(300, 2, "SURVIVAL"), (77, 1, "ESCAPE")])
def test_maze_rollout_mirrors_equal_primary(torch_mod, n, view_grid, task_type):
    torch = torch_mod
    from metagym_b200 import BatchedMetaMaze2D, MazeTaskSampler
    T, w = 40, 2 * view_grid + 1
    fields = {"obs": ((T, n, w, w), torch.float32), "act": ((T, n), torch.int32), "rew": ((T, n), torch.float64),
              "done": ((T, n), torch.uint8)}
    a0, a1, ref = _arenas(torch, fields, 3)
    rs = np.random.RandomState(3)
    tasks = [MazeTaskSampler(n=9, allow_loops=True, crowd_ratio=0.3, rng=rs) for _ in range(5)]

    def run(arena, mirrors):
        env = BatchedMetaMaze2D(max_steps=25, task_type=task_type, view_grid=view_grid, num_envs=n, squeeze=False,
                                auto_reset=True)
        env.set_task(tasks)
        env.reset()
        env.set_mirrors(mirrors)
        env.rollout(T, act_seed=2, out={"obs": arena["obs"], "rew": arena["rew"], "done": arena["done"],
                                        "act": arena["act"]})
        torch.cuda.synchronize()
        env.close()

    run(ref, [])
    run(a0, [a1.buf.data_ptr() - a0.buf.data_ptr()])
    assert torch.equal(a0.buf, ref.buf) and torch.equal(a1.buf, a0.buf)
    assert int(ref["done"].sum()) > 0


def test_mirror_argument_checks(torch_mod):
    from metagym_b200 import BatchedQuadrotor, _lib
    env = BatchedQuadrotor(num_envs=4, device=0, squeeze=False)
    with pytest.raises(_lib.MgbError):
        env.set_mirrors([8])                     # not a multiple of 16 bytes
    with pytest.raises(_lib.MgbError):
        env.set_mirrors([16] * 8)                # more than MGB_MAX_MIRRORS
    env.set_mirrors([])
    env.close()


CHILD = r"""
import ctypes, sys
sys.path.insert(0, %r)
import torch
from metagym_b200 import _lib
from metagym_b200.rollout import _RawDeviceMemory
lib = _lib.load()
handle = (ctypes.c_uint8 * 64)(*bytes.fromhex(sys.argv[1]))
nbytes = int(sys.argv[2])
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
q = ctypes.c_void_p()
_lib.check(lib.mgb_peer_open(0, handle, ctypes.byref(q)))
t = torch.as_tensor(_RawDeviceMemory(q.value, nbytes), device="cuda:0")
assert int(t[:16].sum()) == 16 * 7, "parent's content not visible"
t.copy_((torch.arange(nbytes, device="cuda") %% 251).to(torch.uint8))
torch.cuda.synchronize()
del t
_lib.check(lib.mgb_peer_close(0, q))
print("child ok")
"""


def test_peer_memory_export_open_across_processes(torch_mod):
    """mgb_peer_alloc/export here, mgb_peer_open/close in a child process: the child sees our bytes and we see its."""
    torch = torch_mod
    from metagym_b200 import _lib
    from metagym_b200.rollout import _RawDeviceMemory
    lib = _lib.load()
    nbytes = 1 << 20
    p = ctypes.c_void_p()
    _lib.check(lib.mgb_peer_alloc(0, nbytes, ctypes.byref(p)))
    t = torch.as_tensor(_RawDeviceMemory(p.value, nbytes), device="cuda:0")
    assert int(t.sum()) == 0                     # mgb_peer_alloc zero-fills
    t.fill_(7)
    torch.cuda.synchronize()
    handle = (ctypes.c_uint8 * 64)()
    _lib.check(lib.mgb_peer_export(0, p, handle))
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT, bytes(handle).hex(), str(nbytes)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stdout + r.stderr
    want = (torch.arange(nbytes, device="cuda") % 251).to(torch.uint8)
    assert torch.equal(t, want)
    del t
    _lib.check(lib.mgb_peer_free(0, p))


MC_CHILD = r"""
import os, sys
sys.path.insert(0, %r)
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[1], RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
from metagym_b200 import BatchedQuadrotor, BatchedMetaMaze2D, MazeTaskSampler
from metagym_b200.rollout import MulticastArena, RolloutArena
T, n = 10, 1000
qf = {"obs": ((T, n, 19), torch.float32), "act": ((T, n, 4), torch.float32), "rew": ((T, n), torch.float32),
      "done": ((T, n), torch.uint8)}
mf = {"obs": ((T, n, 5, 5), torch.float32), "act": ((T, n), torch.int32), "rew": ((T, n), torch.float64),
      "done": ((T, n), torch.uint8)}
try:
    qa, ma = MulticastArena(qf, dev), MulticastArena(mf, dev)
except Exception as e:
    print("NO_MULTICAST", repr(e)[:300]); sys.exit(0)
qr, mr = RolloutArena(qf, dev), RolloutArena(mf, dev)
rs = np.random.RandomState(3)
tasks = [MazeTaskSampler(n=9, allow_loops=True, crowd_ratio=0.3, rng=rs) for _ in range(5)]
for arena_q, arena_m, mc in ((qr, mr, False), (qa, ma, True)):
    q = BatchedQuadrotor(task="velocity_control", num_envs=n, device=0, squeeze=False, auto_reset=True, nt=30, dt=0.01)
    m = BatchedMetaMaze2D(max_steps=25, task_type="SURVIVAL", view_grid=2, num_envs=n, squeeze=False, auto_reset=True)
    m.set_task(tasks); q.reset(); m.reset()
    if mc:
        q.set_multicast(arena_q.multicast_delta); m.set_multicast(arena_m.multicast_delta)
    for _ in range(2):
        q.rollout(T, act_seed=5, out={k: arena_q[k] for k in ("obs", "rew", "done", "act")})
        m.rollout(T, act_seed=2, out={k: arena_m[k] for k in ("obs", "rew", "done", "act")})
    torch.cuda.synchronize()
    if mc:
        arena_q.sync(); arena_m.sync()
    q.close(); m.close()
for k in qf:
    assert torch.equal(qa.gathered[k][0], qr[k]), "quad " + k
for k in mf:
    assert torch.equal(ma.gathered[k][0], mr[k]), "maze " + k
assert int(qr["done"].sum()) > 0 and int(mr["done"].sum()) > 0
# misaligned batch sizes are refused, not mis-stored
q = BatchedQuadrotor(num_envs=6, device=0, squeeze=False)
q.reset(); q.set_multicast(qa.multicast_delta)
try:
    q.rollout(2)
    print("MISSING_CHECK")
except Exception as e:
    assert "num_envs %% 4" in str(e), str(e)
print("MC_OK")
dist.destroy_process_group()
"""


def test_multicast_rollout_outputs_equal_plain(torch_mod):
    """XM=2 instantiations (multimem.st through an NVSwitch multicast mapping, world of one rank): the arena the switch
    writes equals the plainly stored outputs bit for bit.  Skipped where the driver offers no multicast object."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-c", MC_CHILD % ROOT, str(port)], capture_output=True, text=True, timeout=600)
    if "NO_MULTICAST" in r.stdout:
        # shown as XFAIL (not a silent skip) on boxes whose driver offers no multicast object for a single GPU; the XM=2
        # instantiations then have 2- and 8-GPU evidence only (profiles/r1_bench_mixed_8gpu.json, byte-for-byte check
        # against an NCCL all-gather inside scripts/bench_mixed.py)
        pytest.xfail("no NVSwitch multicast object on this box: " + r.stdout.strip()[-200:])
    assert r.returncode == 0 and "MC_OK" in r.stdout and "MISSING_CHECK" not in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_mirror_window_refuses_outputs_outside_the_arena(torch_mod):
    """With mirrors on and the arena window registered (arena.attach), a rollout into any other buffer is refused
    instead of storing to `pointer + delta` at a meaningless address; detaching restores plain rollouts."""
    torch = torch_mod
    from metagym_b200 import BatchedMetaMaze2D, BatchedQuadrotor, MazeTaskSampler, _lib
    T, n = 4, 64
    qf = {"obs": ((T, n, 16), torch.float32), "act": ((T, n, 4), torch.float32), "rew": ((T, n), torch.float32),
          "done": ((T, n), torch.uint8)}
    a0, a1 = _arenas(torch, qf, 2)
    env = BatchedQuadrotor(num_envs=n, device=0, squeeze=False, auto_reset=True)
    env.reset()
    delta = a1.buf.data_ptr() - a0.buf.data_ptr()
    env.set_mirrors([delta], window=(a0.buf.data_ptr(), a0.nbytes))
    env.rollout(T, act_seed=1, out=a0.views)                          # inside the window: fine, and mirrored
    torch.cuda.synchronize()
    assert torch.equal(a0.buf, a1.buf)
    with pytest.raises(_lib.MgbError):
        env.rollout(T, act_seed=1)                                    # own buffers: outside -> refused
    with pytest.raises(_lib.MgbError):
        env.rollout(T, act_seed=1, out=a1.views)                      # the other arena: outside -> refused
    with pytest.raises(_lib.MgbError):
        env.rollout(2 * T, act_seed=1, out=a0.views)                  # starts inside but runs past the slot -> refused
    env.set_mirrors([])
    env.rollout(T, act_seed=1)                                        # plain rollout works again
    env.close()
    mf = {"obs": ((T, n, 3, 3), torch.float32), "act": ((T, n), torch.int32), "rew": ((T, n), torch.float64),
          "done": ((T, n), torch.uint8)}
    m0, m1 = _arenas(torch, mf, 2)
    maze = BatchedMetaMaze2D(max_steps=20, task_type="ESCAPE", view_grid=1, num_envs=n, squeeze=False, auto_reset=True)
    maze.set_task(MazeTaskSampler(n=9, rng=np.random.RandomState(0)))
    maze.reset()
    maze.set_mirrors([m1.buf.data_ptr() - m0.buf.data_ptr()], window=(m0.buf.data_ptr(), m0.nbytes))
    maze.rollout(T, act_seed=2, out=m0.views)
    torch.cuda.synchronize()
    assert torch.equal(m0.buf, m1.buf)
    with pytest.raises(_lib.MgbError):
        maze.rollout(T, act_seed=2, want_actions=True)
    with pytest.raises(_lib.MgbError):
        maze.rollout(3 * T, act_seed=2, out=m0.views)                 # extent check, not just the start pointer
    maze.set_mirrors([])
    maze.rollout(T, act_seed=2)
    maze.close()
