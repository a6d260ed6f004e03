"""gym registration ids of the reference, pointing at the batched engine.

The reference registers its envs on import (metagym/quadrotor/__init__.py:20-32, metagym/metamaze/__init__.py:21-54):
    quadrotor-v0, meta-maze-continuous-3D-v0, meta-maze-discrete-3D-v0, meta-maze-2D-v0
`register_envs()` does the same with this package's classes as entry points and the reference's default kwargs, except
that `enable_render` defaults to False (the batched engine is headless; the reference's default True opens a pygame
window).  Extra kwargs (`num_envs`, `device`, `auto_reset`, ...) pass through `gym.make(id, num_envs=4096, ...)`.

gym (or gymnasium) is optional: when neither is importable `register_envs()` returns an empty list and the classes are
used directly.  Called automatically by `import metagym_b200` when gym is importable.
"""

SPECS = [
    ("quadrotor-v0", "metagym_b200.quadrotor:BatchedQuadrotor",
     {"dt": 0.01, "nt": 1000, "seed": 0, "task": "no_collision", "map_file": None, "simulator_conf": None,
      "healthy_reward": 1.0}),
    ("meta-maze-continuous-3D-v0", "metagym_b200.metamaze:BatchedMetaMazeContinuous3D",
     {"enable_render": False, "render_scale": 480, "resolution": (256, 256), "max_steps": 5000, "task_type": "SURVIVAL"}),
    ("meta-maze-discrete-3D-v0", "metagym_b200.metamaze:BatchedMetaMazeDiscrete3D",
     {"enable_render": False, "render_scale": 480, "resolution": (256, 256), "max_steps": 200, "task_type": "SURVIVAL"}),
    ("meta-maze-2D-v0", "metagym_b200.metamaze:BatchedMetaMaze2D",
     {"enable_render": False, "max_steps": 200, "view_grid": 1, "task_type": "SURVIVAL"}),
]


def _registry_module():
    for name in ("gym", "gymnasium"):
        try:
            mod = __import__(name + ".envs.registration", fromlist=["register"])
            return mod
        except Exception:
            continue
    return None


def register_envs(registry=None):
    """Register the four ids; returns the list of ids registered (empty without gym).  `registry`: a module/object with a
    `register(id=..., entry_point=..., kwargs=...)` callable (tests pass a stub)."""
    reg = registry if registry is not None else _registry_module()
    if reg is None:
        return []
    done = []
    for env_id, entry, kwargs in SPECS:
        try:
            reg.register(id=env_id, entry_point=entry, kwargs=dict(kwargs))
            done.append(env_id)
        except Exception:           # already registered (e.g. the reference imported first): leave the existing entry
            continue
    return done
