"""compute-sanitizer workload for maze2d_rollout_kernel: ragged sizes, both store paths, both task types."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metagym_b200 import BatchedMetaMaze2D, MazeTaskSampler
rs = np.random.RandomState(0)
tasks = [MazeTaskSampler(n=11, allow_loops=True, crowd_ratio=0.35, rng=rs) for _ in range(8)]
for tt in ("SURVIVAL", "ESCAPE"):
    for n, g in ((300, 2), (77, 1), (1, 3)):
        env = BatchedMetaMaze2D(max_steps=20, task_type=tt, view_grid=g, num_envs=n, squeeze=False, auto_reset=True)
        env.set_task(tasks); env.reset()
        env.rollout(45, act_seed=1, want_actions=True)
        env.rollout(7, actions=torch.randint(0, 4, (7, n), device="cuda", dtype=torch.int32))
        env.close()
torch.cuda.synchronize()
print("sanitizer workload done")
