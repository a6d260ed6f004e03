"""MazeTaskSampler replays the reference's task sampler (maze_task.py:41-190) sample for sample: fixtures recorded by
tests/golden/gen_maze_tasks.py from the unmodified reference, `random.seed(s); numpy.random.seed(s)` then two tasks."""
import os
import random

import numpy as np
import pytest

from metagym_b200.metamaze import MazeTaskSampler

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
from gen_maze_tasks import CASES  # noqa: E402  (the case table only; the reference loader is not touched)

G = np.load(os.path.join(HERE, "golden", "maze_tasks_golden.npz"))


def check(task, seed, rep):
    pre = "s%d.r%d." % (seed, rep)
    assert np.array_equal(np.asarray(task.cell_walls), G[pre + "walls"]), "walls"
    assert np.array_equal(np.asarray(task.cell_texts), G[pre + "texts"]), "texts"
    assert np.array_equal(np.asarray(task.food_rewards, dtype=np.float64), G[pre + "food"]), "food"
    assert np.array_equal(np.asarray(task.food_interval), G[pre + "interval"]), "interval"
    sc = np.array([task.start[0], task.start[1], task.goal[0], task.goal[1], task.cell_size, task.wall_height,
                   task.agent_height, task.initial_life, task.max_life, task.step_reward, task.goal_reward])
    assert np.array_equal(sc, G[pre + "scalars"]), "scalars"


@pytest.mark.parametrize("seed,kw", CASES)
def test_global_streams_replay_reference(seed, kw):
    random.seed(seed)
    np.random.seed(seed)
    for rep in range(2):
        check(MazeTaskSampler(n_texts=int(G["n_texts"][0]), **kw), seed, rep)


@pytest.mark.parametrize("seed,kw", CASES[:5])
def test_private_streams_replay_reference(seed, kw):
    """seed= / explicit generator objects give the same tasks without touching the global generators."""
    state = random.getstate()
    t0 = MazeTaskSampler(seed=seed, **kw)
    check(t0, seed, 0)
    py, npr = random.Random(seed), np.random.RandomState(seed)
    for rep in range(2):
        check(MazeTaskSampler(py_random=py, np_random=npr, **kw), seed, rep)
    assert random.getstate() == state


def test_distribution_sampler_still_valid():
    rs = np.random.RandomState(0)
    t = MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs)
    w = np.asarray(t.cell_walls)
    assert w[0].all() and w[-1].all() and w[:, 0].all() and w[:, -1].all()
    assert w[1:-1, 1:-1].sum() <= 0.35 * 13 * 13 + 1
