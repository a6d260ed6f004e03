"""Tiny driver for ncu captures: N steps of mgb_quad_step at the bench shape (no graphs, no extras)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metagym_b200 import BatchedQuadrotor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
env = BatchedQuadrotor(task="velocity_control", dt=0.005, nt=1000, seed=list(range(64)), num_envs=n, squeeze=False,
                       auto_reset=True)
env.reset()
G = 8
acts = torch.rand((G, n, 4), device="cuda") * 14.9 + 0.1
obs = torch.empty((G, n, 19), device="cuda"); rew = torch.empty((G, n), device="cuda")
done = torch.empty((G, n), dtype=torch.uint8, device="cuda")
for t in range(steps):
    env.step(acts[t % G], out=(obs[t % G], rew[t % G], done[t % G]))
env.rollout(G, actions=acts, out={"obs": obs, "rew": rew, "done": done, "act": None})
torch.cuda.synchronize()
