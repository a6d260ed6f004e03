"""Generate tests/golden/maze_continuous_golden.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/gen_maze_continuous.py

MetaMazeContinuous3D (metagym/metamaze/envs/maze_continuous_3d.py, dynamics.py) stepped with float32 actions (what
`action_space.sample()` of the reference yields), procedural textures injected as in gen_maze.py.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refload  # noqa: E402
from gen_maze import task_arrays, bfs_path  # noqa: E402
from metagym_b200.textures import synthetic_textures  # noqa: E402


def steer(core, task, targets, rng):
    """float32 (turn, walk) that drives toward the next BFS waypoint of the current target cell, with noise."""
    walls = np.asarray(task.cell_walls)
    cs = task.cell_size
    pos = np.asarray(core._agent_loc, dtype=np.float64)
    cell = (int(pos[0] / cs), int(pos[1] / cs))
    while targets and targets[0] == cell:
        targets.pop(0)
    if not targets:
        return np.array([rng.uniform(-1, 1), rng.uniform(-1, 1)], dtype=np.float32)
    path = bfs_path(walls, cell, targets[0])
    if path is None or len(path) < 2:
        targets.pop(0)
        return np.array([rng.uniform(-1, 1), 0.5], dtype=np.float32)
    wp = (np.array(path[1]) + 0.5) * cs
    want = np.arctan2(wp[1] - pos[1], wp[0] - pos[0])
    err = (want - float(core._agent_ori) + np.pi) % (2 * np.pi) - np.pi
    turn = np.clip(err / (0.1 * 3.1415926), -1.3, 1.3) + rng.normal(0, 0.05)
    walk = (1.2 if abs(err) < 0.4 else 0.15) + rng.normal(0, 0.05)
    return np.array([turn, walk], dtype=np.float32)


def record(ns, task_type, task, n_act, max_steps, resolution, rng):
    env = ns.maze_env.MetaMazeContinuous3D(enable_render=False, resolution=resolution, max_steps=max_steps,
                                           task_type=task_type)
    env.set_task(task)
    rec = dict(rew=[], done=[], pos=[], ori=[], grid=[], steps=[], life=[], obs=[], act=[])
    rec["reset_obs"] = np.asarray(env.reset())

    def new_targets():
        if task_type == "ESCAPE":
            return [tuple(task.goal)]
        t = [tuple(c) for c in np.argwhere(np.asarray(task.food_rewards) > 0.01)]
        rng.shuffle(t)
        return t[:5] + t[:2]

    targets = new_targets()
    for k in range(n_act):
        a = steer(env.maze_core, task, targets, rng)
        if k % 9 == 8:
            a = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.2, 0.2)], dtype=np.float32)   # clip + walking backward
        rec["act"].append(a)
        o, r, d, info = env.step(a)
        core = env.maze_core
        rec["rew"].append(float(r)); rec["done"].append(bool(d))
        rec["pos"].append(np.asarray(core._agent_loc, dtype=np.float32))
        rec["ori"].append(float(core._agent_ori))
        rec["grid"].append([int(core._agent_grid[0]), int(core._agent_grid[1])])
        rec["steps"].append(int(info["steps"]))
        rec["life"].append(float(getattr(core, "_life", 0.0)))
        if k % 3 == 0 or d or r > 0:
            rec["obs"].append(np.asarray(o))
            rec.setdefault("obs_idx", []).append(k)
        if d:
            env.reset()
            targets = new_targets()
    out = {k: np.asarray(v) for k, v in rec.items()}
    assert out["obs"].max() < 32767
    out["obs"] = out["obs"].astype(np.int16)
    out["reset_obs"] = out["reset_obs"].astype(np.int16)
    return out


def main():
    ns = _refload.load_reference()
    grounds, ceil = synthetic_textures(seed=0)
    ns.MAZE_TASK_MANAGER.grounds = grounds.astype(np.float32)
    ns.MAZE_TASK_MANAGER.ceil = ceil.astype(np.uint8)
    out = {}
    cases = [("c3d_surv", "SURVIVAL", dict(n=9, allow_loops=True, crowd_ratio=0.35, food_interval=25, food_density=0.12),
              200, 320, (48, 32)),
             ("c3d_esc", "ESCAPE", dict(n=9, step_reward=-0.01, goal_reward=1.0), 300, 380, (32, 32))]
    for k, (name, tt, skw, max_steps, n_act, res) in enumerate(cases):
        random.seed(50 + k)
        np.random.seed(50 + k)
        task = ns.MazeTaskSampler(**skw)
        rng = np.random.RandomState(200 + k)
        rec = record(ns, tt, task, n_act, max_steps, res, rng)
        for kk, v in task_arrays(task).items():
            out["%s.task.%s" % (name, kk)] = v
        for kk, v in rec.items():
            out["%s.%s" % (name, kk)] = v
        out["%s.meta" % name] = np.array([0 if tt == "SURVIVAL" else 1, max_steps, res[0], res[1]], dtype=np.int32)
        print(name, "steps", n_act, "dones", int(rec["done"].sum()), "reward>0", int((rec["rew"] > 0).sum()),
              "distinct cells", len(set(map(tuple, rec["grid"].tolist()))))
    path = os.path.join(HERE, "maze_continuous_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
