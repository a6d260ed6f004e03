"""Quick device-side timing of the quadrotor kernels (development aid; bench.py is the contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metagym_b200 import BatchedQuadrotor

def timeit(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(iters); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us per iter

CASES = [("velocity_control", 0.005, 65536), ("hovering_control", 0.01, 4096),
         ("velocity_control", 0.005, 4194304), ("hovering_control", 0.01, 65536)]
ONLY = int(sys.argv[1]) if len(sys.argv) > 1 else None      # e.g. `quick_quad_bench.py 4096`: that batch size only
if ONLY is not None:
    CASES = [c for c in CASES if c[2] == ONLY]
for task, dt, N in CASES:
    env = BatchedQuadrotor(task=task, dt=dt, nt=1000, seed=list(range(64)), num_envs=N, squeeze=False,
                           auto_reset=True)
    env.reset()
    T = 64 if N <= 65536 else 8
    acts = torch.rand((T, N, 4), device="cuda") * 14.9 + 0.1
    def steps(k):
        for i in range(k):
            env.step(acts[i % T])
    steps(10)
    us = timeit(steps, 200 if N <= 65536 else 20)
    B = 281 if task == "velocity_control" else 269
    print(json.dumps(dict(kind="step-launches", task=task, N=N, us_per_step=us, steps_per_s=N / us * 1e6,
                          GBps=N * B / us * 1e-3)))
    # CUDA graph of T steps
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        steps(3)
        with torch.cuda.graph(g, stream=s):
            steps(T)
    torch.cuda.synchronize()
    def replay(k):
        for i in range(k): g.replay()
    replay(2)
    us = timeit(replay, 10 if N <= 65536 else 3) / T
    print(json.dumps(dict(kind="step-graph", task=task, N=N, us_per_step=us, steps_per_s=N / us * 1e6,
                          GBps=N * B / us * 1e-3)))
    out = env.rollout(T, actions=acts)
    def roll(k):
        for i in range(k): env.rollout(T, actions=acts, out=out)
    roll(2)
    us = timeit(roll, 10 if N <= 65536 else 3) / T
    print(json.dumps(dict(kind="fused-rollout", task=task, N=N, us_per_step=us, steps_per_s=N / us * 1e6)))
    out2 = {"obs": out["obs"], "rew": out["rew"], "done": out["done"], "act": None}
    def roll2(k):
        for i in range(k): env.rollout(T, actions=None, out=out2)
    roll2(2)
    us = timeit(roll2, 10 if N <= 65536 else 3) / T
    print(json.dumps(dict(kind="fused-rollout-philox", task=task, N=N, us_per_step=us, steps_per_s=N / us * 1e6)))
    env.close()
    del env, acts, out, out2, g
    torch.cuda.empty_cache()

# RK4 variant of the bench shape (config 3 wording)
if ONLY is not None:
    sys.exit(0)
env = BatchedQuadrotor(task="velocity_control", dt=0.005, nt=1000, seed=list(range(64)), num_envs=65536, squeeze=False,
                       auto_reset=True, integrator="rk4")
env.reset()
acts = torch.rand((64, 65536, 4), device="cuda") * 14.9 + 0.1
def steps(k):
    for i in range(k):
        env.step(acts[i % 64])
steps(10)
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    steps(3)
    with torch.cuda.graph(g, stream=s):
        steps(64)
torch.cuda.synchronize()
def replay(k):
    for i in range(k): g.replay()
replay(2)
us = timeit(replay, 10) / 64
print(json.dumps(dict(kind="step-graph-rk4", task="velocity_control", N=65536, us_per_step=us, steps_per_s=65536 / us * 1e6,
                      GBps=65536 * 281 / us * 1e-3)))
