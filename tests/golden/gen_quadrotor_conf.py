"""Generate tests/golden/quadrotor_conf_golden.npz by RUNNING THE UNMODIFIED REFERENCE with a NON-DEFAULT simulator
config (build container only).

    python tests/golden/gen_quadrotor_conf.py

metagym/quadrotor/config.json zeroes several terms of the model (off-diagonal inertia, centre-of-gravity offset, CT[2],
initial velocities).  This config makes all of them non-zero, moves one rotor out of the plane, narrows the voltage
range and changes healthy_reward, so that the terms the default config never exercises are pinned as well.
The JSON text is stored in the fixture; the tests write it to a temp file and pass it as `simulator_conf`.
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _refload  # noqa: E402
from gen_quadrotor import sim_state  # noqa: E402

CONF = {
    "precision": 0.001, "quality": 0.8,
    "inertia": {"xx": 0.0150, "xy": 0.0010, "xz": -0.0005, "yy": 0.0120, "yz": 0.0007, "zz": 0.0260},
    "drag": {"m_xx": 0.060, "m_yy": 0.080, "m_zz": 0.045, "f_xx": 0.10, "f_yy": 0.14, "f_zz": 0.09},
    "gravity_center": {"x": 0.010, "y": -0.005, "z": 0.020},
    "thrust": {"CT": ["1.7e-5", "-2.0e-4", "3.0e-6"], "Mm": "0.012", "Jm": "3.0e-4", "RA": "0.25", "phi": "0.016"},
    "propeller": [{"x": 0.20, "y": 0.17, "z": 0.01}, {"x": -0.18, "y": 0.19, "z": 0.0},
                  {"x": -0.21, "y": -0.18, "z": -0.02}, {"x": 0.18, "y": -0.16, "z": 0.0}],
    "fail": {"velocity": 60.0, "w": 400.0, "range": 500.0},
    "electric": {"min_voltage": 0.5, "max_voltage": 12.0},
    "init_velocity": {"x": 0.5, "y": -0.25, "z": 0.125, "noisy": 1.0},
    "init_angular_velocity": {"x": 0.25, "y": 0.5, "z": -0.75, "noisy": 2.0},
}
HEALTHY = 2.0


def record(ns, conf_path, task, dt, nt, seed, np_seed, T, action_fn, n_episodes):
    env = ns.Quadrotor(task=task, dt=dt, nt=nt, seed=seed, simulator_conf=conf_path, healthy_reward=HEALTHY)
    rng = np.random.RandomState(np_seed + 1000)
    rec = dict(pre_state=[], pre_ct=[], act=[], post_state=[], post_ct=[], obs=[], rew=[], done=[], power=[], ep=[],
               reset_noise=[], reset_obs=[], reset_ct=[])
    np.random.seed(np_seed)
    for ep in range(n_episodes):
        st = np.random.get_state()
        rec["reset_noise"].append(np.random.random(12))
        np.random.set_state(st)
        rec["reset_ct"].append(env.ct)
        rec["reset_obs"].append(env.reset())
        for t in range(T):
            a = action_fn(rng, t)
            rec["pre_state"].append(sim_state(env.simulator)); rec["pre_ct"].append(env.ct)
            o, r, d, _ = env.step(a)
            rec["act"].append(a); rec["post_state"].append(sim_state(env.simulator)); rec["post_ct"].append(env.ct)
            rec["obs"].append(o); rec["rew"].append(float(r)); rec["done"].append(bool(d))
            rec["power"].append(float(env.simulator.power)); rec["ep"].append(ep)
            if d:
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    if task == "velocity_control":
        out["targets"] = np.asarray(env.velocity_targets, dtype=np.float32)
    return out


def main():
    ns = _refload.load_reference()
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(CONF, f)
        path = f.name
    out = {"conf_json": np.array(json.dumps(CONF)), "healthy_reward": np.float64(HEALTHY)}
    uni = lambda rng, t: rng.uniform(0.5, 12.0, 4).astype(np.float32)           # noqa: E731
    wide = lambda rng, t: rng.uniform(-1.0, 14.0, 4).astype(np.float32)         # noqa: E731
    spin = lambda rng, t: np.array([12.0, 0.5, 12.0, 0.5], dtype=np.float32)    # yaw spin-up -> w failure raises? no: |w|<400  # noqa: E731
    runs = {"c_hover": ("hovering_control", 0.01, 1000, 0, 21, 120, uni, 2),
            "c_nocol": ("no_collision", 0.01, 50, 0, 22, 120, wide, 2),
            "c_vel": ("velocity_control", 0.005, 40, 4, 23, 90, uni, 2),
            "c_spin": ("hovering_control", 0.01, 1000, 0, 24, 60, spin, 1)}
    for name, (task, dt, nt, seed, np_seed, T, fn, neps) in runs.items():
        try:
            rec = record(ns, path, task, dt, nt, seed, np_seed, T, fn, neps)
        except Exception as e:                      # a physical failure raises in the reference: keep what ran
            print(name, "raised", repr(e))
            continue
        for k, v in rec.items():
            out["%s.%s" % (name, k)] = v
        out["%s.meta" % name] = np.array([{"hovering_control": 1, "no_collision": 0, "velocity_control": 2}[task], dt, nt,
                                          seed], dtype=np.float64)
        print(name, task, "steps", len(rec["rew"]), "dones", int(rec["done"].sum()))
    os.unlink(path)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadrotor_conf_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
