"""Generate tests/golden/maze_golden.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/gen_maze.py

The reference envs (metagym.metamaze, loaded from /root/reference via _refload.py) are stepped with recorded actions on
tasks drawn from the reference's own MazeTaskSampler.  The texture stack of MAZE_TASK_MANAGER is replaced by the
deterministic procedural set of metagym_b200.textures.synthetic_textures(seed=0) before any rendering, so the fixture
does not embed the reference's image files (the renderer is texture-agnostic; maze_discrete_3d.py:114-117).
"""
import os
import random
import sys
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refload  # noqa: E402
from metagym_b200.textures import synthetic_textures  # noqa: E402


def task_arrays(task):
    return dict(walls=np.asarray(task.cell_walls, dtype=np.int8), texts=np.asarray(task.cell_texts, dtype=np.int8),
                food=np.asarray(task.food_rewards, dtype=np.float64),
                interval=np.asarray(task.food_interval, dtype=np.int32),
                scalars=np.array([task.start[0], task.start[1], task.goal[0], task.goal[1], task.cell_size,
                                  task.wall_height, task.agent_height, task.initial_life, task.max_life,
                                  task.step_reward, task.goal_reward], dtype=np.float64))


def bfs_path(walls, src, dst):
    n = walls.shape[0]
    prev = {tuple(src): None}
    q = deque([tuple(src)])
    while q:
        c = q.popleft()
        if c == tuple(dst):
            break
        for d in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            nx = (c[0] + d[0], c[1] + d[1])
            if 0 <= nx[0] < n and 0 <= nx[1] < n and walls[nx] == 0 and nx not in prev:
                prev[nx] = c
                q.append(nx)
    if tuple(dst) not in prev:
        return None
    path = [tuple(dst)]
    while prev[path[-1]] is not None:
        path.append(prev[path[-1]])
    return path[::-1]


def actions_2d(path):
    amap = {(-1, 0): 0, (1, 0): 1, (0, -1): 2, (0, 1): 3}          # DISCRETE_ACTIONS, maze_env.py:14
    return [amap[(b[0] - a[0], b[1] - a[1])] for a, b in zip(path[:-1], path[1:])]


def actions_3d(path, ori):
    """turn (a=0: -1, a=1: +1) until facing the next cell, then a=3 (forward).  ori: 0 +x, 1 +y, 2 -x, 3 -y."""
    face = {(1, 0): 0, (0, 1): 1, (-1, 0): 2, (0, -1): 3}
    out = []
    for a, b in zip(path[:-1], path[1:]):
        want = face[(b[0] - a[0], b[1] - a[1])]
        while ori != want:
            if (want - ori) % 4 == 3:
                out.append(0)
                ori = (ori - 1) % 4
            else:
                out.append(1)
                ori = (ori + 1) % 4
        out.append(3)
    return out, ori


def plan(task, kind, rng, length, task_type):
    """An action list that walks to food cells / the goal (BFS), padded with random moves (incl. into walls)."""
    walls = np.asarray(task.cell_walls)
    pos, ori, acts = tuple(task.start), 0, []
    targets = [tuple(c) for c in np.argwhere(np.asarray(task.food_rewards) > 0.01)]
    rng.shuffle(targets)
    if task_type == "ESCAPE":
        targets = [tuple(task.goal)]
    else:
        targets = targets[:4] + targets[:1]          # revisit the first food cell (eaten -> respawn timing)
    for tg in targets:
        for _ in range(rng.randint(0, 4)):            # random prefix (may bump into walls / step back)
            acts.append(int(rng.randint(4)))
        # random prefix changes the pose: re-simulate to find it
        pos, ori = simulate(walls, task.start, acts, kind)
        path = bfs_path(walls, pos, tg)
        if path is None:
            continue
        if kind == "2D":
            acts += actions_2d(path)
        else:
            a, ori = actions_3d(path, ori)
            acts += a
        pos, ori = simulate(walls, task.start, acts, kind)
        if len(acts) >= length:
            break
    while len(acts) < length:
        acts.append(int(rng.randint(4)))
    return acts[:length]


def simulate(walls, start, acts, kind):
    n = walls.shape[0]
    x, y, ori = int(start[0]), int(start[1]), 0
    for a in acts:
        if kind == "2D":
            dx, dy = [(-1, 0), (1, 0), (0, -1), (0, 1)][a]
            if walls[x + dx, y + dy] < 1:
                x, y = x + dx, y + dy
        else:
            turn, mv = [(-1, 0), (1, 0), (0, -1), (0, 1)][a]
            ori = (ori + turn) % 4
            dx, dy = [(1, 0), (0, 1), (-1, 0), (0, -1)][ori]
            tx, ty = x + dx * mv, y + dy * mv
            if 0 <= tx < n and 0 <= ty < n and walls[tx, ty] == 0:
                x, y = tx, ty
    return (x, y), ori


def record(ns, kind, task_type, task, acts, max_steps, resolution=None, view_grid=1, keep_obs=None):
    if kind == "2D":
        env = ns.MetaMaze2D(enable_render=False, max_steps=max_steps, task_type=task_type, view_grid=view_grid)
    else:
        env = ns.MetaMazeDiscrete3D(enable_render=False, resolution=resolution, max_steps=max_steps,
                                    task_type=task_type)
    env.set_task(task)
    rec = dict(act=[], rew=[], done=[], agent=[], life=[], obs=[], obs_idx=[], episode=[])
    obs0 = env.reset()
    rec["reset_obs"] = np.asarray(obs0)
    ep = 0
    for t, a in enumerate(acts):
        o, r, d, info = env.step(int(a))
        core = env.maze_core
        ori = getattr(core, "_agent_ori_index", 0)
        rec["act"].append(a)
        rec["rew"].append(float(r))
        rec["done"].append(bool(d))
        rec["agent"].append([int(core._agent_grid[0]), int(core._agent_grid[1]), int(ori), int(info["steps"])])
        rec["life"].append(float(getattr(core, "_life", 0.0)))
        rec["episode"].append(ep)
        if keep_obs is None or t in keep_obs:
            rec["obs"].append(np.asarray(o))
            rec["obs_idx"].append(t)
        if d:
            env.reset()
            ep += 1
    out = {k: np.asarray(v) for k, v in rec.items()}
    if kind == "3D":
        assert out["obs"].max() < 32767
        out["obs"] = out["obs"].astype(np.int16)
        out["reset_obs"] = out["reset_obs"].astype(np.int16)
    return out


def main():
    ns = _refload.load_reference()
    grounds, ceil = synthetic_textures(seed=0)
    ns.MAZE_TASK_MANAGER.grounds = grounds.astype(np.float32)       # same dtypes the reference keeps
    ns.MAZE_TASK_MANAGER.ceil = ceil.astype(np.uint8)
    out = {}

    def sample(k, **kw):
        random.seed(k)
        np.random.seed(k)
        return ns.MazeTaskSampler(**kw)

    cases = [
        # name, kind, task_type, sampler kwargs, max_steps, n_actions, resolution, view_grid, kept frames
        ("m2d_surv", "2D", "SURVIVAL", dict(n=9, food_interval=6, food_density=0.05), 60, 150, None, 1, None),
        ("m2d_surv_g2", "2D", "SURVIVAL", dict(n=15, allow_loops=True, crowd_ratio=0.35), 200, 260, None, 2, None),
        ("m2d_esc", "2D", "ESCAPE", dict(n=15, allow_loops=True, crowd_ratio=0.35), 200, 120, None, 1, None),
        ("m3d_surv", "3D", "SURVIVAL", dict(n=15, allow_loops=True, crowd_ratio=0.35, food_interval=8,
                                            food_density=0.04), 200, 90, (48, 32), 1, None),
        ("m3d_esc", "3D", "ESCAPE", dict(n=9, step_reward=-0.01, goal_reward=1.0), 50, 70, (32, 32), 1, None),
        ("m3d_big", "3D", "SURVIVAL", dict(n=15, allow_loops=True, crowd_ratio=0.35, food_density=0.04), 200, 40,
         (128, 128), 1, {0, 3, 7, 12, 18, 25, 33, 39}),
    ]
    for k, (name, kind, tt, skw, max_steps, n_act, res, g, keep) in enumerate(cases):
        task = sample(k, **skw)
        rng = np.random.RandomState(100 + k)
        acts = plan(task, kind, rng, n_act, tt)
        rec = record(ns, kind, tt, task, acts, max_steps, res, g, keep)
        for kk, v in task_arrays(task).items():
            out["%s.task.%s" % (name, kk)] = v
        for kk, v in rec.items():
            out["%s.%s" % (name, kk)] = v
        out["%s.meta" % name] = np.array([0 if kind == "2D" else 1, 0 if tt == "SURVIVAL" else 1, max_steps, g,
                                          (res or (0, 0))[0], (res or (0, 0))[1]], dtype=np.int32)
        print(name, "steps", len(acts), "dones", int(rec["done"].sum()), "reward>0 steps",
              int((rec["rew"] > 0).sum()), "frames", len(rec["obs"]))

    # 64 tasks of the BASELINE config-4 shape, for the full-size tests and the benchmark (tasks only)
    tw, tt_, tf, ti, tsc = [], [], [], [], []
    for k in range(8):
        task = sample(1000 + k, n=15, allow_loops=True, crowd_ratio=0.35)
        a = task_arrays(task)
        tw.append(a["walls"]); tt_.append(a["texts"]); tf.append(a["food"]); ti.append(a["interval"])
        tsc.append(a["scalars"])
    out["tasks15.walls"], out["tasks15.texts"] = np.asarray(tw), np.asarray(tt_)
    out["tasks15.food"], out["tasks15.interval"], out["tasks15.scalars"] = np.asarray(tf), np.asarray(ti), np.asarray(tsc)

    path = os.path.join(HERE, "maze_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
