"""CPU: the C-ABI library loads and exports exactly the functions include/mgb200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from metagym_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmgb200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and include/mgb200.h disagree"
    assert b"sm_100a" in lib.mgb_version()


def test_struct_layouts_match_header():
    """ctypes mirrors of the ABI structs have the C sizes (checked against a tiny gcc-compiled sizeof program)."""
    import ctypes
    import subprocess
    import tempfile
    from metagym_b200 import _lib
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "mgb200.h"\nint main(){printf("%zu %zu %zu\\n",'
                             'sizeof(mgb_quad_cfg),sizeof(mgb_maze_task_scalars),sizeof(mgb_maze_cfg));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.QuadCfg), ctypes.sizeof(_lib.MazeTaskScalars), ctypes.sizeof(_lib.MazeCfg)]


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (this container has none); with one, this is a no-op."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from metagym_b200 import BatchedQuadrotor, MgbError
    with pytest.raises(MgbError):
        BatchedQuadrotor(num_envs=4)
    from metagym_b200 import _lib
    assert _lib.load().mgb_device_count() < 0 or _lib.load().mgb_device_count() == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the product package (nor a script) may import it; only tests/,
    __graft_entry__.smoke() and bench.py's CPU legs do."""
    for sub in ("metagym_b200", "scripts"):
        d = os.path.join(ROOT, sub)
        for fn in sorted(os.listdir(d)):
            if not fn.endswith(".py"):
                continue
            src = open(os.path.join(d, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), (sub, fn)


def test_host_config_parsing_agrees_with_the_oracle():
    """metagym_b200.quadrotor.build_cfg (what the product hands to mgb_quad_create) and oracle.quad_oracle.make_cfg are
    written independently from QuadrotorSim._parse_cfg (quadrotorsim.py:50-109): same derived numbers for the default
    config and for the non-default one of tests/golden/gen_quadrotor_conf.py; malformed configs raise RuntimeError with
    the reference's prefix."""
    import json
    import numpy as np
    from metagym_b200.quadrotor import build_cfg, velocity_task_actions
    from oracle import quad_oracle as qo
    g = np.load(os.path.join(ROOT, "tests", "golden", "quadrotor_conf_golden.npz"))
    for params in (qo.DEFAULT_PARAMS, json.loads(str(g["conf_json"]))):
        c = build_cfg(params, 0.005, 40, "velocity_control", 2.0)
        o = qo.make_cfg(params)
        f32 = lambda xs: [float(np.float32(x)) for x in xs]      # noqa: E731
        assert f32(c.inv_inertia) == f32(o.Iinv) and f32(c.drag_m) == f32(o.Dm) and f32(c.drag_f) == f32(o.Df)
        assert f32(c.gravity_center) == f32(o.cg) and f32(c.propeller) == f32(o.prop) and f32(c.propeller_norm) == f32(o.lm)
        assert f32(c.ct) == f32([o.ct0, o.ct1, o.ct2]) and (c.mm, c.jm, c.phi, c.ra) == (o.mm, o.jm, o.phi, o.ra)
        assert (c.precision, c.quality) == (o.h, o.m) and (c.min_voltage, c.max_voltage) == (o.vmin, o.vmax)
        assert (c.fail_velocity, c.fail_range, c.fail_w) == (o.fail_v, o.fail_r, o.fail_w)
        assert c.z_offset == 0.0 and c.healthy_reward == 2.0 and c.nt == 40
        a = velocity_task_actions(params, 7, 3)
        assert a.dtype == np.float32 and a.shape == (7, 4) and a.min() >= o.vmin and a.max() <= o.vmax
    assert build_cfg(qo.DEFAULT_PARAMS, 0.01, 10, "hovering_control", 1.0).z_offset == 5.0
    bad = dict(qo.DEFAULT_PARAMS)
    del bad["thrust"]
    with pytest.raises(RuntimeError, match="Error in loading configuration"):
        build_cfg(bad, 0.01, 10, "hovering_control", 1.0)


def test_gym_registration_ids_match_the_reference():
    """quadrotor/__init__.py:20-32 and metamaze/__init__.py:21-54 register four ids on import; register_envs() registers the
    same ids with the same default kwargs (enable_render off: headless) on whatever `register` it is given."""
    from metagym_b200.registration import SPECS, register_envs
    seen = {}

    class Reg(object):
        @staticmethod
        def register(id, entry_point, kwargs):
            seen[id] = (entry_point, kwargs)

    ids = register_envs(Reg)
    assert ids == ["quadrotor-v0", "meta-maze-continuous-3D-v0", "meta-maze-discrete-3D-v0", "meta-maze-2D-v0"]
    assert seen["quadrotor-v0"][1] == {"dt": 0.01, "nt": 1000, "seed": 0, "task": "no_collision", "map_file": None,
                                       "simulator_conf": None, "healthy_reward": 1.0}
    assert seen["meta-maze-discrete-3D-v0"][1]["max_steps"] == 200 and seen["meta-maze-2D-v0"][1]["view_grid"] == 1
    assert seen["meta-maze-continuous-3D-v0"][1]["resolution"] == (256, 256)
    import importlib
    for _, entry, _ in SPECS:                      # every entry point resolves to a class of this package
        mod, cls = entry.split(":")
        assert hasattr(importlib.import_module(mod), cls)
