"""Per-step time of the direct float64 raycaster (many-task regime: one task per env, pose cache off) and of the continuous
maze, eager launches (development aid; bench.py --workload maze3d reports many_tasks from graph replays)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import fast_tasks
from metagym_b200 import BatchedMetaMazeDiscrete3D, BatchedMetaMazeContinuous3D
n = 1024
env = BatchedMetaMazeDiscrete3D(resolution=(128, 128), max_steps=200, num_envs=n, squeeze=False, auto_reset=True, obs_dtype="uint8", cache=False)
env.set_task(fast_tasks(n, seed=5), env2task=np.arange(n)); env.reset()
acts = torch.randint(0, 4, (64, n), device="cuda", dtype=torch.int32)
for t in range(20): env.step(acts[t % 64])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(100): env.step(acts[t % 64])
e1.record(); torch.cuda.synchronize()
print("discrete 3-D direct renderer, 1024 envs 128x128 uint8: %.1f us/step" % (e0.elapsed_time(e1) * 10))
env.close()
cont = BatchedMetaMazeContinuous3D(resolution=(128, 128), max_steps=200, num_envs=n, squeeze=False, auto_reset=True, obs_dtype="uint8")
cont.set_task(fast_tasks(64, seed=5)); cont.reset()
a2 = torch.rand((64, n, 2), device="cuda") * 2 - 1
for t in range(20): cont.step(a2[t % 64])
torch.cuda.synchronize()
e0.record()
for t in range(100): cont.step(a2[t % 64])
e1.record(); torch.cuda.synchronize()
print("continuous 3-D, 1024 envs 128x128 uint8: %.1f us/step" % (e0.elapsed_time(e1) * 10))
cont.close()
