"""Generate tests/golden/maze_geom_golden.npz by RUNNING THE UNMODIFIED REFERENCE with NON-default maze geometry
(build container only).

    python tests/golden/gen_maze_geom.py

Every task of maze_golden.npz has the sampler's default geometry (cell 2.0, wall 3.2, eye 1.6: powers of two almost
everywhere).  Here cell_size / wall_height / agent_height are 1.5 / 2.5 / 0.9 and 3.0 / 4.0 / 2.2, so the general
(division) paths of the renderer and the texture-to-cell ratios that are not powers of two are pinned as well; one
continuous-3-D episode runs on the same geometry.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refload  # noqa: E402
from gen_maze import plan, record, task_arrays  # noqa: E402
import gen_maze_continuous as gc  # noqa: E402
from metagym_b200.textures import synthetic_textures  # noqa: E402


def main():
    ns = _refload.load_reference()
    grounds, ceil = synthetic_textures(seed=0)
    ns.MAZE_TASK_MANAGER.grounds = grounds.astype(np.float32)
    ns.MAZE_TASK_MANAGER.ceil = ceil.astype(np.uint8)
    out = {}
    cases = [
        ("g3d_surv", "3D", "SURVIVAL", dict(n=11, allow_loops=True, crowd_ratio=0.3, cell_size=1.5, wall_height=2.5,
                                            agent_height=0.9, food_density=0.06, food_interval=7), 120, 80, (48, 32)),
        ("g3d_esc", "3D", "ESCAPE", dict(n=9, cell_size=3.0, wall_height=4.0, agent_height=2.2, step_reward=-0.02,
                                         goal_reward=1.5), 60, 70, (40, 24)),
    ]
    for k, (name, kind, tt, skw, max_steps, n_act, res) in enumerate(cases):
        random.seed(70 + k)
        np.random.seed(70 + k)
        task = ns.MazeTaskSampler(**skw)
        rng = np.random.RandomState(300 + k)
        acts = plan(task, kind, rng, n_act, tt)
        rec = record(ns, kind, tt, task, acts, max_steps, res, 1, None)
        for kk, v in task_arrays(task).items():
            out["%s.task.%s" % (name, kk)] = v
        for kk, v in rec.items():
            out["%s.%s" % (name, kk)] = v
        out["%s.meta" % name] = np.array([1, 0 if tt == "SURVIVAL" else 1, max_steps, 1, res[0], res[1]], dtype=np.int32)
        print(name, "steps", len(acts), "dones", int(rec["done"].sum()), "reward>0", int((rec["rew"] > 0).sum()))
    # continuous 3-D on the 1.5 / 2.5 / 0.9 geometry
    random.seed(80)
    np.random.seed(80)
    task = ns.MazeTaskSampler(n=9, allow_loops=True, crowd_ratio=0.3, cell_size=1.5, wall_height=2.5, agent_height=0.9,
                              food_density=0.10, food_interval=20)
    rec = gc.record(ns, "SURVIVAL", task, 200, 150, (40, 32), np.random.RandomState(400))
    for kk, v in task_arrays(task).items():
        out["gc3d.task.%s" % kk] = v
    for kk, v in rec.items():
        out["gc3d.%s" % kk] = v
    out["gc3d.meta"] = np.array([0, 150, 40, 32], dtype=np.int32)
    print("gc3d steps 200 dones", int(rec["done"].sum()), "reward>0", int((rec["rew"] > 0).sum()))
    path = os.path.join(HERE, "maze_geom_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
