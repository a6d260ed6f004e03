"""compute-sanitizer workload for the kernels added or rewritten in round 2 (small shapes, ragged sizes):
quadrotor tile / wide / packed / streaming step kernels (MGB_* select them: run once per setting), host path, fused rollout;
maze3d_step_kernel (ring pipeline, variant frames), two-kernel compose path, direct renderer with the logic kernel ahead,
update_tasks, device task sampler, float32 observations, large-view_grid maze2d.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_r2.py [quad|maze|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metagym_b200 import (BatchedQuadrotor, BatchedMetaMaze2D, BatchedMetaMazeDiscrete3D, BatchedMetaMazeContinuous3D,
                          MazeTaskSampler)

what = sys.argv[1] if len(sys.argv) > 1 else "all"


def quad():
    # 9999 envs: one wave -> quad_step_wide_kernel; 310001 envs: multi-wave -> quad_stream_kernel (TMA double buffer)
    for n in (333, 64, 1, 9999, 310001):
        for task in ("velocity_control", "hovering_control"):
            if n > 1000 and task != "velocity_control":
                continue
            kw = dict(seed=[0, 1, 2], nt=50) if task == "velocity_control" else {}
            env = BatchedQuadrotor(dt=0.005, task=task, num_envs=n, auto_reset=True, squeeze=False, **kw)
            env.reset()
            a = torch.rand((n, 4), device="cuda") * 14.9 + 0.1
            for _ in range(3):
                env.step(a)
            env.step(a.cpu().numpy())                      # host path (staging + copies on the caller's stream)
            env.rollout(4, act_seed=3)
            print("quad", n, task, env.step_kernel_name())
            env.close()
    torch.cuda.synchronize()


def maze():
    rs = np.random.RandomState(0)
    tasks = [MazeTaskSampler(n=9, allow_loops=True, crowd_ratio=0.3, food_density=0.05, rng=rs) for _ in range(3)]
    for fused in ("1", "0"):
        os.environ["MGB_MAZE_FUSED_STEP"] = fused
        for dt in ("uint8", "float32"):
            n = 37
            env = BatchedMetaMazeDiscrete3D(resolution=(32, 32), max_steps=6, num_envs=n, squeeze=False, auto_reset=True,
                                            obs_dtype=dt, task_type="SURVIVAL")
            env.set_task(tasks); env.reset()
            for t in range(8):
                env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.int32))
            env.close()
    os.environ.pop("MGB_MAZE_FUSED_STEP")
    # direct renderer: one task per env, update_tasks + device sampler
    n = 19
    env = BatchedMetaMazeDiscrete3D(resolution=(32, 32), max_steps=5, num_envs=n, squeeze=False, auto_reset=True,
                                    obs_dtype="uint8", cache=False)
    many = [MazeTaskSampler(n=9, allow_loops=True, crowd_ratio=0.3, food_density=0.05, rng=rs) for _ in range(n)]
    env.set_task(many, env2task=np.arange(n)); env.reset()
    for t in range(7):
        _, _, done, _ = env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.int32))
        if t == 2:
            env.update_tasks(np.array([1, 5], dtype=np.int32), [many[0], many[3]])
        env.resample_tasks(mask=done if torch.is_tensor(done) else None, seed=t, food_density=0.05, crowd_ratio=0.3)
    env.close()
    cont = BatchedMetaMazeContinuous3D(resolution=(32, 32), max_steps=5, num_envs=n, squeeze=False, auto_reset=True,
                                       obs_dtype="uint8")
    cont.set_task(tasks); cont.reset()
    for t in range(4):
        cont.step(torch.rand((n, 2), device="cuda") * 2 - 1)
    cont.close()
    e2 = BatchedMetaMaze2D(max_steps=10, view_grid=6, num_envs=45, squeeze=False, auto_reset=True)
    e2.set_task(tasks); e2.reset()
    for t in range(4):
        e2.step(torch.randint(0, 4, (45,), device="cuda", dtype=torch.int32))
    e2.close()
    torch.cuda.synchronize()


if what in ("quad", "all"):
    quad()
if what in ("maze", "all"):
    maze()
print("sanitizer workload done")
