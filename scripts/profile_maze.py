"""Tiny driver for ncu captures: a few maze3d / maze2d steps at the config-4 per-GPU shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metagym_b200 import BatchedMetaMaze2D, BatchedMetaMazeDiscrete3D, MazeTaskSampler
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps3 = int(sys.argv[2]) if len(sys.argv) > 2 else 6      # > 150: steady state (foods eaten, variant frames in use)
rs = np.random.RandomState(0)
tasks = [MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs) for _ in range(64)]
env = BatchedMetaMazeDiscrete3D(resolution=(128, 128), max_steps=200, num_envs=n3, squeeze=False, auto_reset=True,
                                obs_dtype="uint8")
env.set_task(tasks); env.reset()
acts3 = torch.randint(0, 4, (64, n3), device="cuda", dtype=torch.int32)
for t in range(steps3):
    env.step(acts3[t % 64])
e2 = BatchedMetaMaze2D(max_steps=200, task_type="ESCAPE", view_grid=1, num_envs=1048576, squeeze=False, auto_reset=True)
e2.set_task(tasks); e2.reset()
for t in range(6):
    e2.step(torch.randint(0, 4, (1048576,), device="cuda", dtype=torch.int32))
torch.cuda.synchronize()
