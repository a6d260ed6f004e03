"""Loader for the UNMODIFIED reference (PaddlePaddle/MetaGym) used only to GENERATE golden fixtures.

This module is test infrastructure.  It is imported only by the ``gen_*.py`` scripts in this directory and by the
``reference``-marked tests that run when ``/root/reference`` is mounted (the build container).  Nothing on the GPU
box imports it: ``/root/reference`` does not exist there, the committed ``*.npz`` fixtures travel instead.

Shims -- none touches hot-path arithmetic:
  * ``np.int = int``, ``np.product = np.prod``  (removed from numpy >= 1.24 / 2.0; used at
    metagym/quadrotor/quadrotorsim.py:243,250 and metagym/metamaze/envs/maze_task.py:101)
  * a stub ``gym`` module (Env, Space, spaces.Box/Discrete, envs.registration.register, error, utils.seeding)
  * a stub ``pygame`` module whose ``image.load`` is backed by PIL and whose ``surfarray.array3d`` returns the
    (W, H, 3) uint8 x-major array pygame would return.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("METAGYM_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "metagym"))


def _install_numpy_shims():
    if not hasattr(np, "int"):
        np.int = int  # noqa
    if not hasattr(np, "product"):
        np.product = np.prod  # noqa


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")

    class Env(object):
        pass

    class Space(object):
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = dtype

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low)
            self.high = np.asarray(high)
            if shape is None:
                shape = self.low.shape
            Space.__init__(self, shape, dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(np.float32)

    class Discrete(Space):
        def __init__(self, n):
            self.n = n
            Space.__init__(self, (), np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete, spaces.Space = Box, Discrete, Space
    error = types.ModuleType("gym.error")
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registry = {}

    def register(id, entry_point=None, kwargs=None, **kw):
        registry[id] = (entry_point, kwargs or {})

    registration.register = register
    registration.registry = registry
    envs.registration = registration
    gym.Env, gym.Space, gym.spaces, gym.error, gym.utils, gym.envs = Env, Space, spaces, error, utils, envs
    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.error": error, "gym.utils": utils,
        "gym.utils.seeding": seeding, "gym.envs": envs, "gym.envs.registration": registration,
    })


def _install_pygame_stub():
    if "pygame" in sys.modules:
        return
    from PIL import Image

    pygame = types.ModuleType("pygame")

    class _Surface(object):
        def __init__(self, arr):
            self.arr = arr

    def _load(path):
        img = Image.open(path).convert("RGB")
        a = np.asarray(img, dtype=np.uint8)          # (H, W, 3) row-major
        return _Surface(np.ascontiguousarray(a.transpose(1, 0, 2)))  # pygame is x-major: (W, H, 3)

    image = types.ModuleType("pygame.image")
    image.load = _load
    surfarray = types.ModuleType("pygame.surfarray")
    surfarray.array3d = lambda surf: np.array(surf.arr)
    surfarray.make_surface = lambda arr: _Surface(arr)
    font = types.ModuleType("pygame.font")
    font.init = lambda: None
    pygame.image, pygame.surfarray, pygame.font = image, surfarray, font
    pygame.init = lambda: None
    pygame.Surface = _Surface
    sys.modules.update({"pygame": pygame, "pygame.image": image, "pygame.surfarray": surfarray,
                        "pygame.font": font})


_loaded = {}


def load_reference():
    """Import the reference packages; returns a namespace with the symbols the generators need."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_metagym_ref")
    _install_numpy_shims()
    _install_gym_stub()
    _install_pygame_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    ns.quadrotorsim = importlib.import_module("metagym.quadrotor.quadrotorsim")
    ns.quadrotor_env = importlib.import_module("metagym.quadrotor.env")
    ns.QuadrotorSim = ns.quadrotorsim.QuadrotorSim
    ns.Quadrotor = ns.quadrotor_env.Quadrotor
    ns.quad_config = os.path.join(REF_ROOT, "metagym", "quadrotor", "config.json")
    try:
        ns.maze_env = importlib.import_module("metagym.metamaze.envs.maze_env")
        ns.maze_task = importlib.import_module("metagym.metamaze.envs.maze_task")
        ns.ray_caster = importlib.import_module("metagym.metamaze.envs.ray_caster_utils")
        ns.MetaMaze2D = ns.maze_env.MetaMaze2D
        ns.MetaMazeDiscrete3D = ns.maze_env.MetaMazeDiscrete3D
        ns.MAZE_TASK_MANAGER = ns.maze_task.MAZE_TASK_MANAGER
        ns.MazeTaskSampler = ns.maze_task.MazeTaskSampler
        ns.TaskConfig = ns.maze_task.MazeTaskManager.TaskConfig
    except Exception as e:  # pragma: no cover
        ns.maze_error = e
    _loaded["ns"] = ns
    return ns
