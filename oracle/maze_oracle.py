"""TEST INFRASTRUCTURE -- ctypes front end of oracle/maze_oracle.c (CPU restatement of the MetaMaze path).

`OracleMaze` mirrors the call sequence of the reference envs (metagym/metamaze/envs/maze_env.py:16-75,155-206):
set_task(TaskConfig-like) -> reset() -> step(action) -> (obs, reward, done, info).  One instance = one env.
"""
import ctypes

import numpy as np

from . import build as _build

c_i32, c_f64 = ctypes.c_int32, ctypes.c_double
MAX_N = 31


class _Cfg(ctypes.Structure):
    _fields_ = [("n", c_i32), ("task_type", c_i32), ("max_steps", c_i32), ("view_grid", c_i32), ("res_h", c_i32),
                ("res_v", c_i32), ("max_vision", c_f64), ("fov", c_f64), ("l_focal", c_f64), ("text_size", c_f64)]


class _Task(ctypes.Structure):
    _fields_ = [("start", c_i32 * 2), ("goal", c_i32 * 2), ("cell_size", c_f64), ("wall_height", c_f64),
                ("agent_height", c_f64), ("initial_life", c_f64), ("max_life", c_f64), ("step_reward", c_f64),
                ("goal_reward", c_f64)]


class _Env(ctypes.Structure):
    _fields_ = [("gx", c_i32), ("gy", c_i32), ("ori", c_i32), ("steps", c_i32), ("life", c_f64),
                ("cur_food", c_f64 * (MAX_N * MAX_N)), ("revival", c_i32 * (MAX_N * MAX_N)),
                ("wait", c_i32 * (MAX_N * MAX_N))]


class _Cont(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_float * 2), ("pos_is_list", c_i32), ("ori", c_f64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        assert _lib.mo_env_size() == ctypes.sizeof(_Env) and _lib.mo_cont_size() == ctypes.sizeof(_Cont)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleMaze(object):
    def __init__(self, kind, task_type="SURVIVAL", max_steps=200, view_grid=1, resolution=(128, 128), textures=None,
                 max_vision=12.0, fov=0.6 * 3.1415926, l_focal=0.20, text_size=1.0):
        assert kind in ("2D", "3D", "C3D")
        self.cont = _Cont()
        self.kind = kind
        self.cfg = _Cfg()
        self.cfg.task_type = {"SURVIVAL": 0, "ESCAPE": 1}[task_type]
        self.cfg.max_steps = max_steps
        self.cfg.view_grid = view_grid
        self.cfg.res_h, self.cfg.res_v = resolution
        self.cfg.max_vision, self.cfg.fov, self.cfg.l_focal, self.cfg.text_size = max_vision, fov, l_focal, text_size
        if kind in ("3D", "C3D"):
            grounds, ceil = textures
            self.tex = np.ascontiguousarray(grounds, dtype=np.uint8)
            self.ceil = np.ascontiguousarray(ceil, dtype=np.uint8)
            self.ts = self.tex.shape[1]
            assert self.tex.shape[1:] == (self.ts, self.ts, 3) and self.ceil.shape == (self.ts, self.ts, 3)
        self.env = _Env()
        self.need_task = True

    def set_task(self, task):
        n = int(np.shape(task.cell_walls)[0])
        assert n <= MAX_N
        self.cfg.n = n
        self.walls = np.ascontiguousarray(task.cell_walls, dtype=np.int8)
        self.texts = np.ascontiguousarray(task.cell_texts, dtype=np.int8)
        self.food = np.ascontiguousarray(task.food_rewards, dtype=np.float64)
        self.interval = np.ascontiguousarray(task.food_interval, dtype=np.int32)
        t = _Task()
        t.start[:] = [int(task.start[0]), int(task.start[1])]
        t.goal[:] = [int(task.goal[0]), int(task.goal[1])]
        t.cell_size, t.wall_height, t.agent_height = task.cell_size, task.wall_height, task.agent_height
        t.initial_life, t.max_life = task.initial_life, task.max_life
        t.step_reward, t.goal_reward = task.step_reward, task.goal_reward
        self.task = t
        self.need_task = False

    def _observe(self):
        L = lib()
        if self.kind == "2D":
            w = 2 * self.cfg.view_grid + 1
            obs = np.zeros((w, w), dtype=np.float32)
            L.mo_observe_2d(ctypes.byref(self.cfg), ctypes.byref(self.task), ctypes.byref(self.env), _p(self.walls),
                            _p(obs))
            return obs
        H, V = self.cfg.res_h, self.cfg.res_v
        obs = np.zeros((H, V, 3), dtype=np.int32)
        scratch = np.zeros((H, V), dtype=np.float32)
        if self.kind == "C3D":
            L.mo_observe_c3d(ctypes.byref(self.cfg), ctypes.byref(self.task), ctypes.byref(self.env),
                             ctypes.byref(self.cont), _p(self.walls), _p(self.texts), _p(self.tex), _p(self.ceil),
                             ctypes.c_int(self.ts), _p(obs), _p(scratch))
            return obs
        L.mo_observe_3d(ctypes.byref(self.cfg), ctypes.byref(self.task), ctypes.byref(self.env), _p(self.walls),
                        _p(self.texts), _p(self.tex), _p(self.ceil), ctypes.c_int(self.ts), _p(obs), _p(scratch))
        return obs

    def reset(self):
        if self.need_task:
            raise Exception("Must call \"set_task\" before reset")
        if self.kind == "C3D":
            lib().mo_reset_c3d(ctypes.byref(self.cfg), ctypes.byref(self.task), _p(self.food), _p(self.interval),
                               ctypes.byref(self.env), ctypes.byref(self.cont))
            return self._observe()
        lib().mo_reset(ctypes.byref(self.cfg), ctypes.byref(self.task), _p(self.food), _p(self.interval),
                       ctypes.byref(self.env))
        return self._observe()

    def step(self, action, render=True):
        rew = c_f64(0.0)
        done = ctypes.c_int(0)
        if self.kind == "C3D":
            tr, ws = np.float32(action[0]), np.float32(action[1])
            lib().mo_step_c3d(ctypes.byref(self.cfg), ctypes.byref(self.task), _p(self.walls), _p(self.food),
                              _p(self.interval), ctypes.byref(self.env), ctypes.byref(self.cont),
                              ctypes.c_float(float(tr)), ctypes.c_float(float(ws)), ctypes.byref(rew),
                              ctypes.byref(done))
            obs = self._observe() if render else None
            return obs, rew.value, bool(done.value), {"steps": self.env.steps}
        fn = lib().mo_step_2d if self.kind == "2D" else lib().mo_step_3d
        fn(ctypes.byref(self.cfg), ctypes.byref(self.task), _p(self.walls), _p(self.food), _p(self.interval),
           ctypes.byref(self.env), ctypes.c_int(int(action)), ctypes.byref(rew), ctypes.byref(done))
        obs = self._observe() if render else None
        return obs, rew.value, bool(done.value), {"steps": self.env.steps}

    @property
    def agent(self):
        return (self.env.gx, self.env.gy, self.env.ori, self.env.steps)

    @property
    def pose(self):
        return (np.array([self.cont.pos[0], self.cont.pos[1]], dtype=np.float32), float(self.cont.ori))

    @property
    def life(self):
        return self.env.life


def _throughput_worker(args):
    """One process = one oracle env stepping `seconds` of wall clock with uniform actions; returns (steps, wall)."""
    import time
    kind, seed, seconds, warm = args
    from metagym_b200.metamaze import MazeTaskSampler       # host-side sampler only (no CUDA involved)
    from metagym_b200.textures import synthetic_textures
    rs = np.random.RandomState(100 + seed)
    task = MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs)
    if kind == "3D":
        env = OracleMaze("3D", "SURVIVAL", 200, 1, (128, 128), textures=synthetic_textures(seed=0))
    else:
        env = OracleMaze("2D", "ESCAPE", 200, 1)
    env.set_task(task)
    env.reset()
    n = 0
    tw = time.perf_counter()
    while time.perf_counter() - tw < warm:
        if env.step(int(rs.randint(4)))[2]:
            env.reset()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        if env.step(int(rs.randint(4)))[2]:
            env.reset()
        n += 1
    return n, time.perf_counter() - t0


def measure_throughput_detail(kind="3D", seconds=5.0, processes=None, warmup_seconds=0.5):
    """CPU baseline of the maze step (+ render for "3D"): one oracle env per process.
    -> {"value": aggregate env-steps/s, "rates": per-process, "wall_s": slowest process's timed wall}."""
    import multiprocessing as mp
    import os
    cores = processes or len(os.sched_getaffinity(0))
    lib()                                   # build once in the parent, not in every worker
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_throughput_worker, [(kind, k, seconds, warmup_seconds) for k in range(cores)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {"value": total / wall, "rates": [r[0] / r[1] for r in res], "wall_s": wall, "processes": cores}
