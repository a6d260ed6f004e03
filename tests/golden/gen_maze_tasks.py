"""Generate tests/golden/maze_tasks_golden.npz by RUNNING THE UNMODIFIED REFERENCE task sampler (build container only).

    python tests/golden/gen_maze_tasks.py

For every case: `random.seed(s); numpy.random.seed(s)` and then `MazeTaskSampler(**kw)` of
metagym/metamaze/envs/maze_task.py:41-190, twice in a row (the second task continues both random streams), with the
reference's own texture count (n_texts = 7 from its img/ directory).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402
from gen_maze import task_arrays  # noqa: E402

CASES = [
    (1, dict(n=9, step_reward=-0.01, goal_reward=1.0)),
    (2, dict(n=7, allow_loops=False)),
    (3, dict(n=15, allow_loops=True, crowd_ratio=0.35, food_density=0.12, food_interval=25)),
    (4, dict(n=15, allow_loops=True, crowd_ratio=0.0)),
    (5, dict(n=15, allow_loops=False, food_density=0.05)),
    (6, dict(n=21, allow_loops=True, crowd_ratio=0.5, cell_size=1.0, wall_height=2.0, agent_height=1.0)),
    (7, dict(n=11, allow_loops=True, crowd_ratio=0.2, step_reward=-0.02, food_reward=0.3, initial_life=2.0, max_life=3.0)),
    (8, dict(n=31, allow_loops=True, crowd_ratio=0.3)),
    (9, dict()),
]


def main():
    ns = _refload.load_reference()
    out = {"n_texts": np.array([ns.MAZE_TASK_MANAGER.n_texts], dtype=np.int32)}
    for seed, kw in CASES:
        random.seed(seed)
        np.random.seed(seed)
        for rep in range(2):
            task = ns.MazeTaskSampler(**kw)
            for k, v in task_arrays(task).items():
                out["s%d.r%d.%s" % (seed, rep, k)] = v
        print("seed", seed, kw, "interior walls", int(np.asarray(task.cell_walls)[1:-1, 1:-1].sum()))
    path = os.path.join(HERE, "maze_tasks_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
