"""Generate tests/golden/quadrotor_golden.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/gen_quadrotor.py

Every array is an output of metagym.quadrotor (loaded from /root/reference through tests/golden/_refload.py) run
with numpy %s semantics.  Episodes are recorded with the full simulator state BEFORE each step, so a test can replay
them teacher-forced (one step from the recorded state) or free-running (from the recorded reset state).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _refload  # noqa: E402


def sim_state(sim):
    return np.concatenate([
        np.asarray(sim.global_position, dtype=np.float64), np.asarray(sim.global_velocity, dtype=np.float64),
        np.asarray(sim.body_angular_velocity, dtype=np.float64),
        np.asarray(sim.propeller_angular_velocity, dtype=np.float64),
        np.asarray(sim.rotation_matrix, dtype=np.float64).reshape(-1)])


def record_episode(ns, task, dt, nt, seed, np_seed, T, action_fn, n_episodes=1, map_file=None):
    """Run n_episodes (each ends on done or after T steps) on ONE env object, like a user would."""
    env = ns.Quadrotor(task=task, dt=dt, nt=nt, seed=seed, map_file=map_file)
    rng = np.random.RandomState(np_seed + 1000)
    rec = dict(pre_state=[], pre_ct=[], act=[], post_state=[], post_ct=[], obs=[], rew=[], done=[], power=[],
               ep=[], reset_noise=[], reset_obs=[], reset_ct=[])
    np.random.seed(np_seed)
    for ep in range(n_episodes):
        st = np.random.get_state()
        noise = np.random.random(12)
        np.random.set_state(st)
        rec["reset_ct"].append(env.ct)
        o0 = env.reset()
        rec["reset_noise"].append(noise)
        rec["reset_obs"].append(o0)
        for t in range(T):
            a = action_fn(rng, t)
            rec["pre_state"].append(sim_state(env.simulator))
            rec["pre_ct"].append(env.ct)
            o, r, d, info = env.step(a)
            rec["act"].append(a)
            rec["post_state"].append(sim_state(env.simulator))
            rec["post_ct"].append(env.ct)
            rec["obs"].append(o)
            rec["rew"].append(float(r))
            rec["done"].append(bool(d))
            rec["power"].append(float(env.simulator.power))
            rec["ep"].append(ep)
            if d:
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    if task == "velocity_control":
        out["targets"] = np.asarray(env.velocity_targets, dtype=np.float32)
    return out


def main():
    ns = _refload.load_reference()
    out = {}

    # --- KAT 1: never-reset simulator, one step [5,6,7,8] at dt=0.01 (10 substeps); KAT 2: 200 steps of [5,5,5,5]
    sim = ns.QuadrotorSim()
    sim.get_config(ns.quad_config)
    sim.step([5.0, 6.0, 7.0, 8.0], 0.01)
    out["kat1_state"] = sim_state(sim)
    out["kat1_power"] = np.float64(sim.power)
    sim = ns.QuadrotorSim()
    sim.get_config(ns.quad_config)
    traj = []
    for _ in range(200):
        sim.step([5.0, 5.0, 5.0, 5.0], 0.01)
        traj.append(sim_state(sim))
    out["kat2_states"] = np.asarray(traj)

    uni = lambda rng, t: rng.uniform(0.1, 15.0, 4).astype(np.float32)          # noqa: E731
    wide = lambda rng, t: rng.uniform(-2.0, 18.0, 4).astype(np.float32)        # exercises the voltage clamp  # noqa: E731
    fall = lambda rng, t: np.full(4, 0.1, dtype=np.float32)                    # free fall -> floor collision  # noqa: E731

    runs = {
        "hover_a": ("hovering_control", 0.01, 1000, 0, 0, 150, uni, 2),
        "hover_b": ("hovering_control", 0.01, 1000, 0, 1, 150, wide, 1),
        "hover_fall": ("hovering_control", 0.01, 1000, 0, 2, 400, fall, 2),
        "nocol_fall": ("no_collision", 0.01, 1000, 0, 3, 400, fall, 1),
        "nocol_a": ("no_collision", 0.01, 60, 0, 4, 200, uni, 3),               # nt=60 -> time-limit done
        "vel_a": ("velocity_control", 0.005, 40, 0, 5, 100, uni, 3),
        "vel_b": ("velocity_control", 0.005, 40, 7, 6, 100, uni, 2),
        "vel_c": ("velocity_control", 0.01, 25, 3, 7, 100, wide, 2),
    }
    for name, (task, dt, nt, seed, np_seed, T, fn, neps) in runs.items():
        rec = record_episode(ns, task, dt, nt, seed, np_seed, T, fn, neps)
        for k, v in rec.items():
            out["%s.%s" % (name, k)] = v
        out["%s.meta" % name] = np.array([{"hovering_control": 1, "no_collision": 0, "velocity_control": 2}[task],
                                          dt, nt, seed], dtype=np.float64)
        print(name, task, "steps", len(rec["rew"]), "dones", int(rec["done"].sum()))

    # --- obstacle map (env.py:248-260,293-305): a 12x12 map whose start cell is ringed by obstacle cells, so the window
    # swept by a step almost always contains one and the collision threshold becomes z + 5 < 1 instead of < 0
    import tempfile
    obst = np.zeros((12, 12), dtype=np.int64)
    obst[3:8, 3:8] = 10
    obst[5, 5] = -1
    obst[5, 4] = 0
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        for row in obst:
            f.write(" ".join(str(int(v)).zfill(2) for v in row) + "\n")
        map_path = f.name
    out["map_obst"] = obst.astype(np.int32)
    glide = lambda rng, t: np.array([4.4, 4.6, 4.4, 4.6], dtype=np.float32) + rng.uniform(-0.3, 0.3, 4).astype(np.float32)  # noqa: E731
    for name, (task, np_seed, fn, neps) in {"map_hover": ("hovering_control", 11, glide, 3),
                                            "map_nocol": ("no_collision", 12, fall, 2)}.items():
        rec = record_episode(ns, task, 0.01, 1000, 0, np_seed, 400, fn, neps, map_file=map_path)
        for k, v in rec.items():
            out["%s.%s" % (name, k)] = v
        out["%s.meta" % name] = np.array([{"hovering_control": 1, "no_collision": 0}[task], 0.01, 1000, 0], dtype=np.float64)
        zs = rec["obs"][:, 15]
        print(name, "steps", len(rec["rew"]), "dones", int(rec["done"].sum()), "z at done", zs[rec["done"]])
    os.unlink(map_path)

    # --- velocity-task generator alone (quadrotorsim.py:306-319): seeds 0..5, nt=40, dt=0.005
    tabs = []
    for seed in range(6):
        sim = ns.QuadrotorSim()
        sim.get_config(ns.quad_config)
        tabs.append(np.asarray(sim.define_velocity_control_task(0.005, 40, seed), dtype=np.float32))
    out["veltask_tables"] = np.asarray(tabs)

    out["numpy_version"] = np.array(np.__version__)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadrotor_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
