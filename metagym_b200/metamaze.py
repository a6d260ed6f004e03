"""Batched drop-ins for metagym.metamaze.MetaMaze2D / MetaMazeDiscrete3D
(reference: metagym/metamaze/envs/maze_env.py:16-75, 155-206) plus a host-side task sampler.

Same constructor kwargs, `set_task` / `reset` / `step` protocol and error messages as the reference, with a leading
batch axis over `num_envs` independent instances stepped by libmgb200 (metagym_b200/csrc/maze.cu).  A task is the
reference's `TaskConfig` namedtuple (metagym/metamaze/envs/maze_task.py:15-17); objects produced by the reference's own
`MazeTaskSampler` are accepted as they are (duck-typed by field name).
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import _lib
from .spaces import Box, Discrete
from .textures import synthetic_textures

PI = 3.1415926                                       # metagym/metamaze/envs/dynamics.py:6
DISCRETE_ACTIONS = [(-1, 0), (1, 0), (0, -1), (0, 1)]   # maze_env.py:14

TaskConfig = namedtuple("TaskConfig", ["start", "goal", "cell_walls", "cell_texts", "cell_size", "wall_height",
                                       "agent_height", "initial_life", "max_life", "step_reward", "goal_reward",
                                       "food_rewards", "food_interval"])


class _DSU(object):
    def __init__(self, n):
        self.p = list(range(n))

    def find(self, x):
        while self.p[x] != x:
            self.p[x] = self.p[self.p[x]]
            x = self.p[x]
        return x

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a == b:
            return False
        self.p[a] = b
        return True


def _carve_like_reference(n, allow_loops, crowd_ratio, py):
    """The wall-removal process of maze_task.py:84-151, restated on arrays, consuming `py` (a `random.Random` or the
    `random` module) call for call like the reference: one `shuffle` of the walls still standing per round (in their
    row-major order of creation), then walls are examined in shuffled order until one (a) joins different regions
    without closing a loop, (b) joins regions and closes a loop (only with allow_loops), or (c) once a single region is
    left, wins a `random() < 0.2` draw.  A round that runs out of walls takes the last one examined (reference quirk:
    the loop variables simply survive the for statement); a chosen wall with no open neighbour is skipped.
    Rounds continue while more than one region exists or (allow_loops and interior walls exceed crowd_ratio)."""
    walls = np.ones((n, n), dtype=np.int32)
    walls[1:n:2, 1:n:2] = 0
    region = np.full((n, n), -1, dtype=np.int64)
    rooms = [(i, j) for i in range(1, n - 1) for j in range(1, n - 1) if walls[i, j] == 0]
    for k, (i, j) in enumerate(rooms):
        region[i, j] = k
    members = {k: [c] for k, c in enumerate(rooms)}
    standing = [(i, j) for i in range(1, n - 1) for j in range(1, n - 1) if walls[i, j] > 0]
    limit = (n - 2) * (n - 2) * crowd_ratio
    while len(members) > 1 or (allow_loops and len(standing) > limit):
        order = list(standing)
        py.shuffle(order)
        if not order:
            raise RuntimeError("maze sampler: no wall left to remove")
        for i, j in order:
            ids = [int(region[a, b]) for a, b in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1)) if walls[a, b] < 1]
            keep = min(ids) if ids else -1
            others = set(ids) - {keep}
            repeats = max([ids.count(v) for v in ids] + [1])
            if others and repeats < 2:
                break
            if others and repeats > 1 and allow_loops:
                break
            if allow_loops and len(members) < 2 and py.random() < 0.2:
                break
        if keep < 0:
            continue
        walls[i, j] = 0
        region[i, j] = keep
        members[keep].append((i, j))
        standing.remove((i, j))
        for o in others:
            for a, b in members[o]:
                region[a, b] = keep
            members[keep].extend(members.pop(o))
    return walls


def _sample_task_reference_streams(n, allow_loops, cell_size, wall_height, agent_height, step_reward, goal_reward,
                                   food_reward, initial_life, max_life, food_density, food_interval, crowd_ratio,
                                   n_texts, py, npr):
    """maze_task.py:41-190 with the reference's exact random-stream usage: with py = `random` and npr = `numpy.random`
    after `random.seed(s); numpy.random.seed(s)` (or `random.Random(s)` / `numpy.random.RandomState(s)`) the task is
    the reference's task for that seed, array for array (tests/test_maze_sampler.py, fixtures recorded from the
    unmodified reference)."""
    texts = npr.randint(1, n_texts, size=(n, n))
    m = (n - 1) // 2
    sx = py.randint(0, m - 1) * 2 + 1
    sy = py.randint(0, m - 1) * 2 + 1
    goal = (n - 2, n - 2)
    for _ in range(m):                 # the reference's `break` leaves the inner loop only: m rows of up to m draws,
        for _ in range(m):             # the LAST far-enough candidate of a row that found one wins
            ex = py.randint(0, m - 1) * 2 + 1
            ey = py.randint(0, m - 1) * 2 + 1
            if np.sqrt((ex - sx) ** 2 + (ey - sy) ** 2) > 0.45 * n:
                goal = (ex, ey)
                break
    walls = _carve_like_reference(n, allow_loops, crowd_ratio, py)
    inner = texts[1:-1, 1:-1]
    inner[walls[1:-1, 1:-1] < 1] = 0
    def_goal_reward = -np.sqrt(n) * n * step_reward if goal_reward is None else goal_reward
    assert def_goal_reward > 0, "goal reward must be > 0"
    food = np.clip(npr.rand(n, n) * food_reward, 0.10, food_reward)
    food *= 1.0 - walls
    expected = (n - 1) * (n - 1) * food_density
    while np.sum(food) > expected:
        food *= (npr.rand(n, n) < 0.90).astype("float32")
    interval = food_interval * (food > 1.0e-3).astype("int32")
    return TaskConfig(start=(sx, sy), goal=goal, cell_walls=walls, cell_texts=texts, cell_size=cell_size,
                      step_reward=step_reward, goal_reward=def_goal_reward, wall_height=wall_height,
                      agent_height=agent_height, initial_life=initial_life, max_life=max_life, food_rewards=food,
                      food_interval=interval)


def MazeTaskSampler(n=15, allow_loops=True, cell_size=2.0, wall_height=3.2, agent_height=1.6, step_reward=-0.01,
                    goal_reward=None, food_reward=0.50, initial_life=1.0, max_life=2.0, food_density=0.010,
                    food_interval=100, crowd_ratio=0.0, n_texts=7, rng=None, seed=None, py_random=None, np_random=None):
    """Random maze task with the reference sampler's schema, defaults and constraints (maze_task.py:41-190).

    Default (no `rng`): the reference's procedure replayed on the reference's random streams - the global `random` and
    `numpy.random` modules, or `py_random` / `np_random` instances, or both seeded from `seed` - so that
    `random.seed(s); numpy.random.seed(s); MazeTaskSampler(...)` returns exactly the task the reference returns.

    With `rng` (a numpy RandomState): a faster sampler of the same DISTRIBUTION family (rooms on odd coordinates, a
    random spanning tree by Kruskal over the room lattice, then with `allow_loops` interior walls are knocked out until
    at most `crowd_ratio` of the interior is wall); not sample-identical to the reference.
    """
    assert n > 6, "Minimum required cells are 7"
    assert n % 2 != 0, "Cell Numbers can only be odd"
    assert step_reward < 0, "step_reward must be < 0"
    if rng is None:
        import random as _pyrandom
        py = py_random if py_random is not None else (_pyrandom.Random(seed) if seed is not None else _pyrandom)
        npr = np_random if np_random is not None else (np.random.RandomState(seed) if seed is not None else np.random)
        return _sample_task_reference_streams(n, allow_loops, cell_size, wall_height, agent_height, step_reward,
                                              goal_reward, food_reward, initial_life, max_life, food_density,
                                              food_interval, crowd_ratio, n_texts, py, npr)
    rs = rng
    walls = np.ones((n, n), dtype=np.int32)
    walls[1:n:2, 1:n:2] = 0
    m = (n - 1) // 2
    dsu = _DSU(m * m)
    edges = []
    for a in range(m):
        for b in range(m):
            if a + 1 < m:
                edges.append((2 * a + 2, 2 * b + 1, a * m + b, (a + 1) * m + b))
            if b + 1 < m:
                edges.append((2 * a + 1, 2 * b + 2, a * m + b, a * m + b + 1))
    for k in rs.permutation(len(edges)):
        i, j, u, v = edges[k]
        if dsu.union(u, v):
            walls[i, j] = 0
    if allow_loops:
        interior = walls[1:-1, 1:-1]
        budget = interior.size * crowd_ratio
        cand = [(i, j) for i in range(1, n - 1) for j in range(1, n - 1) if walls[i, j] > 0]
        for k in rs.permutation(len(cand)):
            if interior.sum() <= budget:
                break
            i, j = cand[k]
            if walls[i - 1, j] == 0 or walls[i + 1, j] == 0 or walls[i, j - 1] == 0 or walls[i, j + 1] == 0:
                walls[i, j] = 0
    texts = rs.randint(1, n_texts, size=(n, n))
    texts[walls < 1] = 0
    start = (int(rs.randint(0, m)) * 2 + 1, int(rs.randint(0, m)) * 2 + 1)
    goal = (n - 2, n - 2)
    for _ in range(m * m):
        cand = (int(rs.randint(0, m)) * 2 + 1, int(rs.randint(0, m)) * 2 + 1)
        if np.hypot(cand[0] - start[0], cand[1] - start[1]) > 0.45 * n:
            goal = cand
            break
    def_goal_reward = -np.sqrt(n) * n * step_reward if goal_reward is None else goal_reward
    assert def_goal_reward > 0, "goal reward must be > 0"
    food = np.clip(rs.rand(n, n) * food_reward, 0.10, food_reward) * (1.0 - walls)
    expected = (n - 1) * (n - 1) * food_density
    while food.sum() > expected:
        food *= (rs.rand(n, n) < 0.90).astype("float32")
    interval = food_interval * (food > 1.0e-3).astype("int32")
    return TaskConfig(start=start, goal=goal, cell_walls=walls, cell_texts=texts, cell_size=cell_size,
                      step_reward=step_reward, goal_reward=def_goal_reward, wall_height=wall_height,
                      agent_height=agent_height, initial_life=initial_life, max_life=max_life, food_rewards=food,
                      food_interval=interval)


class _BatchedMazeBase(object):
    KIND = None

    def _setup(self, num_envs, device, task_type, max_steps, auto_reset, env_index_base, squeeze):
        import torch
        assert task_type in ("SURVIVAL", "ESCAPE")
        self._torch = torch
        self.num_envs = int(num_envs)
        self.task_type = task_type
        self.max_steps = max_steps
        self.auto_reset = bool(auto_reset)
        self.env_index_base = int(env_index_base)
        self._squeeze = bool(squeeze) and self.num_envs == 1
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.MgbError("metagym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.load()
        self._h = None
        self._n_cells = None
        self.action_space = Discrete(4)
        self.need_reset = True                         # maze_env.py:41-42
        self.need_set_task = True
        self._rew = torch.empty((self.num_envs,), dtype=torch.float64, device=self.device)
        self._done = torch.empty((self.num_envs,), dtype=torch.uint8, device=self.device)
        self._own_ptrs = None

    def _stream(self):
        return _lib.current_stream(self._torch, self.device)

    def _make_cfg(self, n_cells):
        raise NotImplementedError

    def _after_create(self):
        pass

    def _create(self, n_cells):
        if self._h is not None:
            if n_cells == self._n_cells:
                return
            self._lib.mgb_maze_destroy(self._h)
            self._h = None
        cfg = self._make_cfg(n_cells)
        h = ctypes.c_void_p()
        _lib.check(self._lib.mgb_maze_create(ctypes.byref(h), self.num_envs, ctypes.byref(cfg), self.device.index,
                                             self.env_index_base))
        self._h, self._n_cells = h, n_cells
        _lib.check(self._lib.mgb_maze_set_options(self._h, int(self.auto_reset)))
        self._after_create()

    def set_task(self, task_config, env2task=None):
        """MazeBase.set_task (maze_base.py:19-38).  `task_config`: one TaskConfig (all envs) or a sequence of them;
        env i then runs task env2task[i] (default: (env_index_base + i) % n_tasks)."""
        tasks = [task_config] if hasattr(task_config, "cell_walls") else list(task_config)
        n = int(np.shape(tasks[0].cell_walls)[0])
        for t in tasks:
            w = np.asarray(t.cell_walls)
            assert t.agent_height < t.wall_height and t.agent_height > 0, \
                "the agent height must be > 0 and < wall height"
            assert w.shape == np.shape(t.cell_texts), "the dimension of walls must be equal to textures"
            assert w.shape[0] == w.shape[1], "only support square shape"
            assert w.shape[0] == n, "all tasks of one batch must share the maze size"
        self._create(n)
        K = len(tasks)
        walls = np.ascontiguousarray(np.stack([np.asarray(t.cell_walls) for t in tasks]).astype(np.int8))
        texts = np.ascontiguousarray(np.stack([np.asarray(t.cell_texts) for t in tasks]).astype(np.int8))
        food = np.ascontiguousarray(np.stack([np.asarray(t.food_rewards, dtype=np.float64) for t in tasks]))
        itv = np.ascontiguousarray(np.stack([np.asarray(t.food_interval) for t in tasks]).astype(np.int32))
        sc = (_lib.MazeTaskScalars * K)()
        for k, t in enumerate(tasks):
            sc[k].start[:] = [int(t.start[0]), int(t.start[1])]
            sc[k].goal[:] = [int(t.goal[0]), int(t.goal[1])]
            sc[k].cell_size, sc[k].wall_height, sc[k].agent_height = t.cell_size, t.wall_height, t.agent_height
            sc[k].initial_life, sc[k].max_life = t.initial_life, t.max_life
            sc[k].step_reward, sc[k].goal_reward = t.step_reward, t.goal_reward
        if env2task is None:
            env2task = (np.arange(self.num_envs, dtype=np.int64) + self.env_index_base) % K
        e2t = np.ascontiguousarray(np.asarray(env2task, dtype=np.int32))
        assert e2t.shape == (self.num_envs,)
        self._torch.cuda.synchronize(self.device)
        _lib.check(self._lib.mgb_maze_set_task(self._h, K, walls.ctypes.data, texts.ctypes.data, food.ctypes.data,
                                               itv.ctypes.data, sc, e2t.ctypes.data))
        self.tasks, self.env2task = tasks, e2t
        self.need_set_task = False
        self.need_reset = True

    def update_tasks(self, task_slots, task_configs):
        """Per-episode task resampling: replace entries `task_slots` of the task table by `task_configs`, stream-ordered
        and without a device synchronisation (mgb_maze_update_tasks).  Envs flying a replaced slot restart on the new task.
        With set_task(tasks) of one task per env (env2task = arange), `update_tasks(env_ids, new_tasks)` re-tasks exactly
        those envs.  Needs the direct renderer (cache=False / more tasks than the pose-cache budget)."""
        slots = np.ascontiguousarray(np.asarray(task_slots, dtype=np.int32).reshape(-1))
        tasks = [task_configs] if hasattr(task_configs, "cell_walls") else list(task_configs)
        K = len(tasks)
        assert slots.shape == (K,), "one task per slot"
        walls = np.ascontiguousarray(np.stack([np.asarray(t.cell_walls) for t in tasks]).astype(np.int8))
        texts = np.ascontiguousarray(np.stack([np.asarray(t.cell_texts) for t in tasks]).astype(np.int8))
        food = np.ascontiguousarray(np.stack([np.asarray(t.food_rewards, dtype=np.float64) for t in tasks]))
        itv = np.ascontiguousarray(np.stack([np.asarray(t.food_interval) for t in tasks]).astype(np.int32))
        assert walls.shape[1:] == (self._n_cells, self._n_cells), "all tasks of one batch must share the maze size"
        sc = (_lib.MazeTaskScalars * K)()
        for k, t in enumerate(tasks):
            assert t.agent_height < t.wall_height and t.agent_height > 0, "the agent height must be > 0 and < wall height"
            sc[k].start[:] = [int(t.start[0]), int(t.start[1])]
            sc[k].goal[:] = [int(t.goal[0]), int(t.goal[1])]
            sc[k].cell_size, sc[k].wall_height, sc[k].agent_height = t.cell_size, t.wall_height, t.agent_height
            sc[k].initial_life, sc[k].max_life = t.initial_life, t.max_life
            sc[k].step_reward, sc[k].goal_reward = t.step_reward, t.goal_reward
        _lib.check(self._lib.mgb_maze_update_tasks(self._h, K, slots.ctypes.data, walls.ctypes.data, texts.ctypes.data,
                                                   food.ctypes.data, itv.ctypes.data, sc, self._stream()))
        for k, sl in enumerate(slots):
            self.tasks[int(sl)] = tasks[k]

    def resample_tasks(self, mask=None, seed=0, allow_loops=True, cell_size=2.0, wall_height=3.2, agent_height=1.6,
                       step_reward=-0.01, goal_reward=None, food_reward=0.50, initial_life=1.0, max_life=2.0,
                       food_density=0.010, food_interval=100, crowd_ratio=0.0, n_texts=7):
        """Per-episode task resampling on the device (mgb_maze_resample_tasks): every env with mask[e] != 0 (None: all)
        gets a freshly drawn maze (MazeTaskSampler's keyword arguments and distribution family; counter-based draws keyed
        by (seed, global env index, resample count)) and restarts on it.  One stream-ordered kernel; typical use:
        `obs, rew, done, _ = env.step(a); env.resample_tasks(done)`.  Needs set_task() with one table slot per env
        (env2task = arange) and the direct renderer (cache=False)."""
        m = None
        if mask is not None:
            m = self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        cfg = _lib.MazeSamplerCfg()
        cfg.allow_loops, cfg.n_texts, cfg.food_interval = int(bool(allow_loops)), int(n_texts), int(food_interval)
        cfg.cell_size, cfg.wall_height, cfg.agent_height = cell_size, wall_height, agent_height
        cfg.step_reward, cfg.goal_reward = step_reward, (0.0 if goal_reward is None else goal_reward)
        cfg.food_reward, cfg.initial_life, cfg.max_life = food_reward, initial_life, max_life
        cfg.food_density, cfg.crowd_ratio = food_density, crowd_ratio
        _lib.check(self._lib.mgb_maze_resample_tasks(self._h, _lib.ptr(m), ctypes.byref(cfg), int(seed), self._stream()))
        self.need_reset = False

    def get_tasks(self, task_slots):
        """Tasks currently in the table slots `task_slots` (synchronous read-back) -> list of TaskConfig."""
        slots = np.ascontiguousarray(np.asarray(task_slots, dtype=np.int32).reshape(-1))
        K, n = int(slots.size), self._n_cells
        walls = np.empty((K, n, n), np.int8); texts = np.empty((K, n, n), np.int8)
        food = np.empty((K, n, n), np.float64); itv = np.empty((K, n, n), np.int32)
        sc = (_lib.MazeTaskScalars * K)()
        _lib.check(self._lib.mgb_maze_get_tasks(self._h, K, slots.ctypes.data, walls.ctypes.data, texts.ctypes.data,
                                                food.ctypes.data, itv.ctypes.data, sc))
        return [TaskConfig(start=(int(sc[k].start[0]), int(sc[k].start[1])), goal=(int(sc[k].goal[0]), int(sc[k].goal[1])),
                           cell_walls=walls[k].astype(np.int32), cell_texts=texts[k].astype(np.int64),
                           cell_size=sc[k].cell_size, step_reward=sc[k].step_reward, goal_reward=sc[k].goal_reward,
                           wall_height=sc[k].wall_height, agent_height=sc[k].agent_height,
                           initial_life=sc[k].initial_life, max_life=sc[k].max_life, food_rewards=food[k],
                           food_interval=itv[k]) for k in range(K)]

    def sample_task(self, **kwargs):
        """Convenience: draw one task with the host sampler (reference usage: MazeTaskSampler(...), test.py:12)."""
        return MazeTaskSampler(**kwargs)

    def _out(self, t):
        return t[0] if self._squeeze else t

    def reset(self, mask=None):
        if self.need_set_task:
            raise Exception("Must call \"set_task\" before reset")                  # maze_env.py:49-50
        m = None
        if mask is not None:
            m = self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        _lib.check(self._lib.mgb_maze_reset(self._h, _lib.ptr(m), self._obs.data_ptr(), self._stream()))
        self.need_reset = False
        return self._out(self._obs)

    def step(self, action=None):
        if self.need_reset:
            raise Exception("Must \"reset\" before doing any actions")              # maze_env.py:60-61
        if action is None:
            raise NotImplementedError("keyboard control (action=None) is display-only in the reference")
        torch = self._torch
        if not (hasattr(action, "is_cuda") and action.is_cuda):
            action = torch.as_tensor(np.asarray(action).reshape(self.num_envs), device=self.device)
        act = action
        if act.dtype is not torch.int32 or not act.is_contiguous() or act.numel() != self.num_envs:
            act = action.to(torch.int32).reshape(self.num_envs).contiguous()
        if self._own_ptrs is None or self._own_ptrs[0] != self._obs.data_ptr():
            self._own_ptrs = (self._obs.data_ptr(), self._rew.data_ptr(), self._done.data_ptr())
            self._done_bool = self._done.view(torch.bool)
        p = self._own_ptrs
        rc = self._lib.mgb_maze_step(self._h, act.data_ptr(), p[0], p[1], p[2], self._stream())
        if rc:
            _lib.check(rc)
        info = _LazySteps(self)
        return self._out(self._obs), self._out(self._rew), self._out(self._done_bool), info

    def _rollout(self, T, actions, act_seed, want_actions, out):
        if self.need_reset:
            raise Exception("Must \"reset\" before doing any actions")
        torch = self._torch
        N, dev = self.num_envs, self.device
        if out is None:
            out = {"obs": torch.empty((T, N) + tuple(self._obs.shape[1:]), dtype=self._obs.dtype, device=dev),
                   "rew": torch.empty((T, N), dtype=torch.float64, device=dev),
                   "done": torch.empty((T, N), dtype=torch.uint8, device=dev),
                   "act": torch.empty((T, N), dtype=torch.int32, device=dev) if want_actions else None}
        a = None if actions is None else actions.to(torch.int32).reshape(T, N).contiguous()
        _lib.check(self._lib.mgb_maze_rollout(self._h, int(T), _lib.ptr(a), int(act_seed), _lib.ptr(out.get("act")),
                                              _lib.ptr(out.get("obs")), _lib.ptr(out.get("rew")),
                                              _lib.ptr(out.get("done")), self._stream()))
        return out

    def agent_state(self):
        """-> (agent [N,4] int32 = grid_x, grid_y, ori_index, steps ; life [N] float64)."""
        torch = self._torch
        ag = torch.empty((self.num_envs, 4), dtype=torch.int32, device=self.device)
        life = torch.empty((self.num_envs,), dtype=torch.float64, device=self.device)
        _lib.check(self._lib.mgb_maze_state(self._h, ag.data_ptr(), life.data_ptr(), self._stream()))
        return ag, life

    @property
    def launch_count(self):
        return int(self._lib.mgb_maze_launch_count(self._h)) if self._h else 0

    def render(self, mode="human"):
        raise NotImplementedError("the pygame god-view (maze_base.py:100-189) is out of scope for the batched engine")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mgb_maze_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _LazySteps(dict):
    """info = {"steps": n} (maze_env.py:73): fetched from the device only when somebody looks."""

    def __init__(self, env):
        dict.__init__(self)
        self._env = env

    def __missing__(self, k):
        if k != "steps":
            raise KeyError(k)
        ag, _ = self._env.agent_state()
        v = ag[:, 3]
        self[k] = v[0] if self._env._squeeze else v
        return self[k]

    def __contains__(self, k):
        return k == "steps"


class BatchedMetaMaze2D(_BatchedMazeBase):
    """MetaMaze2D(enable_render, render_scale, max_steps, task_type, view_grid) x num_envs (maze_env.py:155-172)."""
    KIND = 0

    def __init__(self, enable_render=False, render_scale=480, max_steps=5000, task_type="SURVIVAL", view_grid=2,
                 num_envs=1, device=0, auto_reset=False, env_index_base=0, squeeze=True):
        if enable_render:
            raise NotImplementedError("enable_render=True needs a display; the batched engine is headless")
        self.enable_render = False
        self.view_grid = int(view_grid)
        self._setup(num_envs, device, task_type, max_steps, auto_reset, env_index_base, squeeze)
        w = 2 * self.view_grid + 1
        # the reference declares Box(-1, 1, (3,3), int32) but returns float32 (2g+1)^2 arrays (maze_2d.py:92)
        self.observation_space = Box(low=-1, high=1, shape=(w, w), dtype=np.float32)
        self._obs = self._torch.empty((self.num_envs, w, w), dtype=self._torch.float32, device=self.device)

    def _make_cfg(self, n_cells):
        cfg = _lib.MazeCfg()
        cfg.kind, cfg.task_type = 0, {"SURVIVAL": 0, "ESCAPE": 1}[self.task_type]
        cfg.n_cells, cfg.max_steps, cfg.view_grid = n_cells, self.max_steps, self.view_grid
        return cfg

    def _set_window(self, window):
        """window = (base address, bytes) of this rank's arena slot: rollouts writing elsewhere are refused while
        mirrors are on (arena.attach(env) passes it)."""
        base, nbytes = (0, 0) if window is None else (int(window[0]), int(window[1]))
        _lib.check(self._lib.mgb_maze_set_mirror_window(self._h, base, nbytes))

    def set_mirrors(self, byte_deltas, window=None):
        """Every output of rollout() is also stored at `pointer + delta` (see rollout.PeerArena)."""
        self._set_window(window)
        d = np.ascontiguousarray(np.asarray(list(byte_deltas), dtype=np.int64))
        _lib.check(self._lib.mgb_maze_set_mirrors(self._h, int(d.size), _lib.ptr(d) if d.size else None))

    def set_multicast(self, byte_delta, window=None):
        """rollout() outputs go through an NVSwitch multicast mapping at `pointer + byte_delta` (rollout.MulticastArena)."""
        self._set_window(window)
        _lib.check(self._lib.mgb_maze_set_multicast(self._h, int(byte_delta)))

    def rollout(self, T, actions=None, act_seed=0, want_actions=False, out=None):
        """T steps in one launch (mgb_maze_rollout).  actions: [T,N] int32 CUDA tensor or None (device-drawn uniform
        {0..3}).  Returns dict(obs [T,N,<obs of one env>], rew [T,N] f64, done [T,N] u8, act [T,N] i32 or None)."""
        return self._rollout(T, actions, act_seed, want_actions, out)


class BatchedMetaMazeDiscrete3D(_BatchedMazeBase):
    """MetaMazeDiscrete3D(enable_render, render_scale, resolution, max_steps, task_type) x num_envs
    (maze_env.py:16-42).  obs_dtype: 'int32' = exact reference values (what the reference's array holds), 'float32' = the
    same values in the dtype the reference's observation_space declares (maze_env.py:37-39), 'uint8' = min(value, 255).
    textures: (grounds uint8 [n_tex,64,64,3], ceil uint8 [64,64,3]); default = procedural set."""
    KIND = 1

    def __init__(self, enable_render=False, render_scale=480, resolution=(320, 320), max_steps=5000,
                 task_type="SURVIVAL", num_envs=1, device=0, auto_reset=False, env_index_base=0, squeeze=True,
                 obs_dtype="int32", textures=None, max_vision_range=12.0, fol_angle=0.6 * PI, cache=None):
        if enable_render:
            raise NotImplementedError("enable_render=True needs a display; the batched engine is headless")
        self.enable_render = False
        self.resolution = (int(resolution[0]), int(resolution[1]))
        assert obs_dtype in ("int32", "uint8", "float32")
        self.obs_dtype = obs_dtype
        self.max_vision_range, self.fol_angle = max_vision_range, fol_angle
        self.cache = cache                  # None: library default (on, MGB_MAZE_CACHE); False: direct renderer only
        self.textures = textures if textures is not None else synthetic_textures(seed=0)
        self._setup(num_envs, device, task_type, max_steps, auto_reset, env_index_base, squeeze)
        torch = self._torch
        h, v = self.resolution
        self.observation_space = Box(low=0, high=256, shape=(h, v, 3), dtype=np.float32)     # maze_env.py:37-39
        self._obs = torch.empty((self.num_envs, h, v, 3), device=self.device,
                                dtype={"int32": torch.int32, "uint8": torch.uint8, "float32": torch.float32}[obs_dtype])

    def _make_cfg(self, n_cells):
        cfg = _lib.MazeCfg()
        cfg.kind, cfg.task_type = 1, {"SURVIVAL": 0, "ESCAPE": 1}[self.task_type]
        cfg.n_cells, cfg.max_steps, cfg.view_grid = n_cells, self.max_steps, 0
        cfg.res_h, cfg.res_v = self.resolution
        cfg.obs_dtype = {"uint8": 0, "int32": 1, "float32": 2}[self.obs_dtype]
        cfg.max_vision, cfg.fov = self.max_vision_range, self.fol_angle       # maze_discrete_3d.py:22-23
        cfg.l_focal, cfg.text_size = 0.20, 1.0                                # maze_discrete_3d.py:116
        return cfg

    def rollout(self, T, actions=None, act_seed=0, want_actions=False, out=None):
        """T steps in one launch on the pose cache: obs [T,N,res_h,res_v,3] (uint8 or int32), rew, done, act as for
        BatchedMetaMaze2D.rollout."""
        if self.KIND != 1:
            raise NotImplementedError("fused rollout: MetaMaze2D and MetaMazeDiscrete3D")
        return self._rollout(T, actions, act_seed, want_actions, out)

    def cache_info(self):
        """Pose-cache statistics (valid after the first reset()/step()): dict(poses, variant_frames, variant_bits, bytes,
        poses_by_food_count [k = 0..7, >= 8], in_use)."""
        out = (ctypes.c_int64 * 16)()
        _lib.check(self._lib.mgb_maze_cache_info(self._h, out))
        return {"poses": int(out[0]), "variant_frames": int(out[1]), "variant_bits": int(out[2]), "bytes": int(out[3]),
                "poses_by_food_count": [int(out[4 + k]) for k in range(9)], "in_use": bool(out[13])}

    def _after_create(self):
        if self.cache is not None:
            _lib.check(self._lib.mgb_maze_set_cache(self._h, int(bool(self.cache))))
        grounds = np.ascontiguousarray(self.textures[0], dtype=np.uint8)
        ceil = np.ascontiguousarray(self.textures[1], dtype=np.uint8)
        ts = grounds.shape[1]
        assert grounds.shape[1:] == (ts, ts, 3) and ceil.shape == (ts, ts, 3)
        _lib.check(self._lib.mgb_maze_set_textures(self._h, grounds.ctypes.data, grounds.shape[0], ceil.ctypes.data,
                                                   ts))


class BatchedMetaMazeContinuous3D(BatchedMetaMazeDiscrete3D):
    """MetaMazeContinuous3D(enable_render, render_scale, resolution, max_steps, task_type) x num_envs
    (maze_env.py:85-153).  Actions [N, 2] float32 = (turn_rate, walk_speed) in [-1, 1] (clipped like the reference,
    maze_continuous_3d.py:48-49); float32 position / float64 heading exactly as the reference computes them for
    float32 actions.  Every pose is unique, so this env always uses the direct float64 renderer."""
    KIND = 2

    def __init__(self, *args, **kwargs):
        BatchedMetaMazeDiscrete3D.__init__(self, *args, **kwargs)
        self.action_space = Box(low=np.array([-1.0, -1.0]), high=np.array([1.0, 1.0]), dtype=np.float32)

    def _make_cfg(self, n_cells):
        cfg = BatchedMetaMazeDiscrete3D._make_cfg(self, n_cells)
        cfg.kind = 2
        return cfg

    def step(self, action=None):
        if self.need_reset:
            raise Exception("Must \"reset\" before doing any actions")              # maze_env.py:130-131
        if action is None:
            raise NotImplementedError("keyboard control (action=None) is display-only in the reference")
        torch = self._torch
        if not (hasattr(action, "is_cuda") and action.is_cuda):
            action = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(self.num_envs, 2), device=self.device)
        act = action.to(torch.float32).reshape(self.num_envs, 2).contiguous()
        _lib.check(self._lib.mgb_maze_step_continuous(self._h, act.data_ptr(), self._obs.data_ptr(),
                                                      self._rew.data_ptr(), self._done.data_ptr(), self._stream()))
        info = _LazySteps(self)
        return self._out(self._obs), self._out(self._rew), self._out(self._done.view(torch.bool)), info

    def pose(self):
        """-> (pos [N,2] float32 = _agent_loc, ori [N] float64 = _agent_ori)."""
        torch = self._torch
        pos = torch.empty((self.num_envs, 2), dtype=torch.float32, device=self.device)
        ori = torch.empty((self.num_envs,), dtype=torch.float64, device=self.device)
        _lib.check(self._lib.mgb_maze_pose(self._h, pos.data_ptr(), ori.data_ptr(), self._stream()))
        return pos, ori


MetaMaze2D = BatchedMetaMaze2D
MetaMazeDiscrete3D = BatchedMetaMazeDiscrete3D
MetaMazeContinuous3D = BatchedMetaMazeContinuous3D

