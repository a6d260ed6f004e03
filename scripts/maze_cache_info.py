import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import reference_tasks
from metagym_b200 import BatchedMetaMazeDiscrete3D
tasks = reference_tasks(64)
for bits, gb in ((4, 24), (6, 100), (5, 100)):
    os.environ["MGB_MAZE_VARIANT_BITS"] = str(bits); os.environ["MGB_MAZE_CACHE_GB"] = str(gb)
    env = BatchedMetaMazeDiscrete3D(resolution=(128, 128), max_steps=200, num_envs=1024, squeeze=False, auto_reset=True, obs_dtype="uint8")
    t0 = time.time(); env.set_task(tasks); env.reset(); torch.cuda.synchronize(); t1 = time.time()
    print("bits", bits, "budget", gb, "build s %.2f" % (t1 - t0), env.cache_info())
    acts = torch.randint(0, 4, (64, 1024), device="cuda", dtype=torch.int32)
    for t in range(300): env.step(acts[t % 64])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(400): env.step(acts[t % 64])
    e1.record(); torch.cuda.synchronize()
    print("   eager us/step %.2f" % (e0.elapsed_time(e1) * 1e3 / 400))
    env.close()
