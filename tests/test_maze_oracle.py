"""CPU: pin the maze oracle (oracle/maze_oracle.c) against golden vectors recorded from the unmodified reference.
Everything here is integer / exact: grid state, done flags, float64 rewards and life, float32 2-D observations and the
int32 raycast images must match bit for bit."""
import numpy as np
import pytest

from oracle.maze_oracle import OracleMaze
from metagym_b200.textures import synthetic_textures
from util import MAZE_CASES, maze_case


def replay(case, make_env):
    env = make_env()
    env.set_task(case["task"])
    obs0 = env.reset()
    yield ("reset", -1, obs0, None, None, None)
    kept = {int(t): k for k, t in enumerate(case["obs_idx"])}
    for t, a in enumerate(case["act"]):
        obs, rew, done, info = env.step(int(a))
        yield ("step", t, obs, rew, done, (env.agent, env.life, kept.get(t)))
        if done:
            env.reset()


@pytest.fixture(scope="module")
def geom_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "maze_geom_golden.npz"))


GEOM_CASES = ["g3d_surv", "g3d_esc"]      # non-default cell / wall / eye heights (tests/golden/gen_maze_geom.py)


@pytest.mark.parametrize("name", MAZE_CASES + GEOM_CASES)
def test_oracle_matches_reference_episode(maze_golden, geom_golden, name):
    c = maze_case(geom_golden if name in GEOM_CASES else maze_golden, name)
    tex = synthetic_textures(seed=0)

    def make():
        return OracleMaze(c["kind"], c["task_type"], c["max_steps"], c["view_grid"], c["resolution"], textures=tex)

    n_frames = 0
    for what, t, obs, rew, done, extra in replay(c, make):
        if what == "reset":
            assert np.array_equal(np.asarray(obs), c["reset_obs"].astype(obs.dtype))
            assert obs.dtype == (np.float32 if c["kind"] == "2D" else np.int32)
            continue
        agent, life, k = extra
        assert rew == c["rew"][t], (t, rew, c["rew"][t])
        assert done == bool(c["done"][t]), t
        assert tuple(agent) == tuple(int(x) for x in c["agent"][t]), t
        if c["task_type"] == "SURVIVAL":
            assert life == c["life"][t], t
        if k is not None:
            ref = c["obs"][k]
            if c["kind"] == "2D":
                assert obs.dtype == np.float32 and np.array_equal(obs, ref), t
            else:
                assert np.array_equal(obs, ref.astype(np.int32)), (t, int((obs != ref).sum()))
            n_frames += 1
    assert n_frames == len(c["obs_idx"])


def test_values_can_exceed_uint8(maze_golden):
    """Near-floor pixels are lit with v_screen / l_focal > 1 (ray_caster_utils.py:99,114): with a bright ground texture
    the reference's int32 image exceeds 255, which is why the engine offers an exact int32 mode next to the clamped
    uint8 one (MGB_OBS_I32 / MGB_OBS_U8)."""
    c = maze_case(maze_golden, "m3d_big")
    grounds, ceil = synthetic_textures(seed=0)
    grounds = grounds.copy()
    grounds[0] = 255
    env = OracleMaze("3D", "ESCAPE", 200, 1, (128, 128), textures=(grounds, ceil))
    env.set_task(c["task"])
    obs = env.reset()
    assert 300 < int(obs.max()) < 400


@pytest.mark.parametrize("name", ["c3d_surv", "c3d_esc", "gc3d"])
def test_continuous_maze_oracle_matches_reference(cont_golden, geom_golden, name):
    """MetaMazeContinuous3D: float32 positions, float64 headings, rewards, dones and every
    recorded frame of the reference episodes, bit for bit (numba/numpy typing of dynamics.py reproduced in C)."""
    from util import cont_case
    c = cont_case(geom_golden if name == "gc3d" else cont_golden, name)
    tex = synthetic_textures(seed=0)
    env = OracleMaze("C3D", c["task_type"], c["max_steps"], 1, c["resolution"], textures=tex)
    env.set_task(c["task"])
    assert np.array_equal(env.reset(), c["reset_obs"].astype(np.int32))
    kept = {int(t): k for k, t in enumerate(c["obs_idx"])}
    for t, a in enumerate(c["act"]):
        obs, rew, done, info = env.step(a)
        pos, ori = env.pose
        assert np.array_equal(pos, c["pos"][t]) and ori == c["ori"][t], t
        assert rew == c["rew"][t] and done == bool(c["done"][t]) and info["steps"] == int(c["steps"][t]), t
        assert tuple(env.agent[:2]) == tuple(int(x) for x in c["grid"][t])
        if c["task_type"] == "SURVIVAL":
            assert env.life == c["life"][t]
        if t in kept:
            assert np.array_equal(obs, c["obs"][kept[t]].astype(np.int32)), t
        if done:
            env.reset()
    assert c["done"].sum() >= 1 and (c["rew"] > 0).sum() >= 1


@pytest.mark.reference
def test_oracle_vs_reference_real_textures():
    """Build container only: the reference renderer with its own PNG textures vs the oracle, random poses."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refload
    if not _refload.reference_available():
        pytest.skip("reference tree not mounted")
    import random
    ns = _refload.load_reference()
    from metagym_b200.textures import load_texture_dir
    tex = load_texture_dir(os.path.join(_refload.REF_ROOT, "metagym", "metamaze", "envs", "img"))
    ns.MAZE_TASK_MANAGER.grounds = tex[0].astype(np.float32)
    ns.MAZE_TASK_MANAGER.ceil = tex[1]
    random.seed(5)
    np.random.seed(5)
    task = ns.MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, food_density=0.05)
    ref = ns.MetaMazeDiscrete3D(enable_render=False, resolution=(128, 128), max_steps=500, task_type="SURVIVAL")
    ref.set_task(task)
    ora = OracleMaze("3D", "SURVIVAL", 500, 1, (128, 128), textures=tex)
    ora.set_task(task)
    assert np.array_equal(ref.reset(), ora.reset())
    rng = np.random.RandomState(0)
    for t in range(60):
        a = int(rng.randint(4))
        o1, r1, d1, _ = ref.step(a)
        o2, r2, d2, _ = ora.step(a)
        assert np.array_equal(o1, o2), (t, int((o1 != o2).sum()))
        assert r1 == r2 and d1 == d2
