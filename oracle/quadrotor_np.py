"""TEST INFRASTRUCTURE -- single-env numpy port of the reference quadrotor, used as the CPU BASELINE.

BASELINE.json's metric is quoted "vs reference numpy CPU".  The reference itself is pure Python and cannot travel to the
GPU box, so `bench.py --impl reference` (and the `cpu_baseline` leg) time this port instead: one Python object per env,
small numpy arrays, the same numpy calls per substep (matmul / cross / norm / inv on 3-vectors and 3x3 matrices), hence
the same cost profile as metagym/quadrotor/quadrotorsim.py:122-221 + env.py:127-165.  tests/test_quadrotor_oracle.py
checks it against the golden vectors recorded from the unmodified reference (it agrees to the last bit on this
container's numpy because it issues the same numpy operations on the same dtypes).

`substep()` deliberately issues the SAME numpy operations in the SAME order on the SAME dtypes as
quadrotorsim.py:122-208 (a statement-for-statement restatement with renamed variables and pre-bound constants): that is
what makes it bit-identical and cost-identical to the reference, and it is why this file is test infrastructure only --
nothing under metagym_b200/ imports it.
"""
import math
from math import ceil, floor

import numpy as np

from .quad_oracle import DEFAULT_PARAMS

G = np.array([0.0, 0.0, -9.80], dtype=np.float32)


class SimParams(object):
    """Numbers of one simulator config, typed the way the reference holds them (python floats + float32 arrays)."""

    def __init__(self, p=None):
        p = DEFAULT_PARAMS if p is None else p
        self.h = float(p["precision"])
        self.mass = float(p["quality"])
        i = p["inertia"]
        inertia = np.zeros((3, 3)).astype(np.float32)
        for (r, c), k in {(0, 0): "xx", (0, 1): "xy", (0, 2): "xz", (1, 0): "xy", (1, 1): "yy", (1, 2): "yz",
                          (2, 0): "xz", (2, 1): "yz", (2, 2): "zz"}.items():
            inertia[r, c] = float(i[k])
        self.inv_inertia = np.linalg.inv(inertia)
        d = p["drag"]
        self.drag_m = np.diag([float(d["m_xx"]), float(d["m_yy"]), float(d["m_zz"])]).astype(np.float32)
        self.drag_f = np.diag([float(d["f_xx"]), float(d["f_yy"]), float(d["f_zz"])]).astype(np.float32)
        self.cg = np.array([float(p["gravity_center"][k]) for k in "xyz"], dtype=np.float32)
        t = p["thrust"]
        self.ct = [float(x) for x in t["CT"]]
        self.mm, self.jm, self.phi, self.ra = float(t["Mm"]), float(t["Jm"]), float(t["phi"]), float(t["RA"])
        self.fail_v, self.fail_r, self.fail_w = (float(p["fail"][k]) for k in ("velocity", "range", "w"))
        self.arms = np.array([[float(q[k]) for k in "xyz"] for q in p["propeller"]], dtype=np.float32)
        self.vmin, self.vmax = float(p["electric"]["min_voltage"]), float(p["electric"]["max_voltage"])
        self.init_v, self.init_w = p["init_velocity"], p["init_angular_velocity"]


class SimState(object):
    __slots__ = ("p", "v", "w", "rotor", "R", "Rinv", "power")

    def __init__(self):
        z3 = lambda: np.array([0.0] * 3).astype(np.float32)      # noqa: E731
        self.p, self.v, self.w = z3(), z3(), z3()
        self.rotor = np.array([0.0] * 4).astype(np.float32)
        self.R = np.eye(3).astype(np.float32)
        self.Rinv = np.linalg.inv(self.R)
        self.power = 0.0

    def as_row(self):
        return np.concatenate([np.asarray(self.p, np.float64), np.asarray(self.v, np.float64),
                               np.asarray(self.w, np.float64), np.asarray(self.rotor, np.float64),
                               np.asarray(self.R, np.float64).reshape(-1)])


def randomise(st, q):
    """reset(): zero state + signed uniform noise on both velocity vectors (12 global-RNG draws)."""
    fresh = SimState()
    for name in SimState.__slots__:
        setattr(st, name, getattr(fresh, name))
    for attr, spec in (("v", q.init_v), ("w", q.init_w)):
        sign = ((np.random.random(3) > 0.5).astype(int) * 2) - 1.0
        noisy = float(spec["noisy"]) * np.random.random(3)
        setattr(st, attr, np.array([spec["x"], spec["y"], spec["z"]], dtype=np.float32) + noisy * sign)
    return st


def substep(st, q, act):
    """One precision-sized step of the rigid body (rotors -> wrench -> semi-implicit Euler -> matrix inverse)."""
    force = np.zeros(3).astype(np.float32)
    torque = np.zeros(3).astype(np.float32)
    watts = np.zeros(4).astype(np.float32)
    me = np.zeros(4).astype(np.float32)
    for i in range(4):
        volt = act[i]
        if volt > q.vmax:
            volt = q.vmax
        elif volt < q.vmin:
            volt = q.vmin
        back_emf = q.phi * st.rotor[i]
        me[i] = q.phi / q.ra * (volt - back_emf)
        watts[i] = abs(me[i] / q.phi * volt)
        accel = 1.0 / q.jm * (me[i] - q.mm)
        w_new = st.rotor[i] + q.h * accel
        reach = np.linalg.norm(q.arms[i])
        v_body = np.matmul(st.Rinv, st.v)
        sweep = np.cross(st.w, q.arms[i]) * reach
        inflow = v_body[2] + sweep[2]
        sgn = 1.0 if inflow > 0 else -1.0
        lift = q.ct[0] * w_new * w_new + q.ct[1] * w_new * inflow + q.ct[2] * inflow * inflow * sgn
        st.rotor[i] = w_new
        force[2] += lift
        torque += np.cross(-np.array([0.0, 0.0, lift], dtype=np.float32), q.arms[i])
    torque[2] += -me[0] + me[1] - me[2] + me[3]
    f_drag = -np.linalg.norm(st.v) * np.matmul(np.matmul(q.drag_f, st.Rinv), st.v)
    t_drag = -np.linalg.norm(st.w) * np.matmul(q.drag_m, st.w)
    f_grav = np.matmul(st.Rinv, G) * q.mass
    t_grav = -np.cross(f_grav, q.cg)
    f_all = force + f_grav + f_drag
    t_all = torque + t_grav + t_drag
    acc = np.matmul(st.R, f_all / q.mass)
    st.p += st.v * q.h + 0.5 * q.h * q.h * acc
    st.v += q.h * acc
    st.power = np.sum(watts)
    ang_acc = np.matmul(q.inv_inertia, t_all)
    mid = st.w + 0.5 * q.h * ang_acc
    skew = np.zeros((3, 3)).astype(np.float32)
    skew[0, 1], skew[0, 2] = -mid[2], mid[1]
    skew[1, 0], skew[1, 2] = mid[2], -mid[0]
    skew[2, 0], skew[2, 1] = -mid[1], mid[0]
    st.R += q.h * np.matmul(st.R, skew)
    st.w += q.h * ang_acc
    st.Rinv = np.linalg.inv(st.R)
    if np.linalg.norm(st.p) > q.fail_r:
        raise Exception("The quadrotor exists the valid zone")
    if np.linalg.norm(st.v) > q.fail_v:
        raise Exception("The quadrotor has too large velocity to recover")
    if np.linalg.norm(st.w) > q.fail_w:
        raise Exception("The quadrotor has too large angular velocity")


def advance(st, q, act, dt):
    for _ in range(int(dt / q.h)):
        substep(st, q, act)


def velocity_table(q, dt, nt, seed):
    """Velocity targets of the velocity_control task: seeded random actions flown from the zero state."""
    np.random.seed(seed)
    st = SimState()
    rows = []
    for _ in range(nt):
        a = np.random.uniform(low=q.vmin, high=q.vmax, size=4).astype(np.float32)
        advance(st, q, a, dt)
        rows.append(list(st.v))
    return rows


class NumpyQuadrotorEnv(object):
    """gym-style env over the functions above (reset/step like the reference's Quadrotor)."""

    def __init__(self, dt=0.01, nt=1000, seed=0, task="no_collision", healthy_reward=1.0, params=None, map_matrix=None):
        assert task in ("velocity_control", "no_collision", "hovering_control"), "Invalid task setting"
        self.q = SimParams(params)
        self.dt, self.nt, self.task, self.healthy = dt, nt, task, healthy_reward
        self.ct = 0
        self.st = SimState()
        self.z_off = 0
        self.z0 = np.float32(0.0)
        if task == "velocity_control":
            self.targets = velocity_table(self.q, dt, nt, seed)
        else:
            self.z_off = 5.0
            m = np.zeros([100, 100], dtype=np.int32) if map_matrix is None else np.array(map_matrix)
            if map_matrix is None:
                m[50, 50] = -1
            ys, xs = np.where(m == -1)
            assert len(ys) == 1
            self.y_off, self.x_off = ys[0], xs[0]
            m[self.y_off, self.x_off] = 0
            self.map = m

    def _observe(self):
        st = self.st
        bv = np.matmul(st.Rinv, st.v)
        bp = np.matmul(st.Rinv, st.p)
        acc = np.zeros(3, dtype=np.float32) + np.matmul(st.Rinv, G)
        roll = np.arctan2(st.R[2, 1], st.R[2, 2])
        pitch = np.arctan2(-st.R[2, 0], np.sqrt(st.R[2, 1] ** 2 + st.R[2, 2] ** 2))
        yaw = np.arctan2(st.R[1, 0], st.R[0, 0])
        self._bv = bv
        vals = [bv[0], bv[1], bv[2], bp[0], bp[1], bp[2], acc[0], acc[1], acc[2], st.w[0], st.w[1], st.w[2],
                pitch, roll, yaw, st.p[2] + self.z_off]
        if self.task == "velocity_control":
            vals.extend(self.targets[min(self.ct, self.nt - 1)])
        return np.array(vals, dtype=np.float32)

    def reset(self):
        randomise(self.st, self.q)
        self.z0 = np.copy(self.st.p)[2]
        return self._observe()

    def step(self, action):
        self.ct += 1
        cmd = np.asarray(action, np.float32)
        z_before = self.st.p[2] + self.z_off
        if self.task != "velocity_control":
            xy_before = (self.st.p[0] + self.x_off, self.st.p[1] + self.y_off)
        advance(self.st, self.q, cmd.tolist(), self.dt)
        obs = self._observe()
        z_after = self.st.p[2] + self.z_off
        st = self.st
        reward = -min(self.dt * st.power, self.healthy)
        done = False
        if self.task == "velocity_control":
            tgt = np.matmul(st.Rinv, self.targets[self.ct - 1])
            reward += -0.001 * (abs(tgt[0] - self._bv[0]) + abs(tgt[1] - self._bv[1]) + abs(tgt[2] - self._bv[2]))
        else:
            z_lo = int(floor(min(z_before, z_after)))
            z_hi = int(ceil(max(z_before, z_after)))
            xy_after = (st.p[0] + self.x_off, st.p[1] + self.y_off)
            x_lo, x_hi = int(floor(min(xy_before[0], xy_after[0]))), int(ceil(max(xy_before[0], xy_after[0])))
            y_lo, y_hi = int(floor(min(xy_before[1], xy_after[1]))), int(ceil(max(xy_before[1], xy_after[1])))
            swept = self.map[y_lo:y_hi + 1, x_lo:x_hi + 1]
            hit = z_lo < np.any(swept) or z_hi < np.any(swept)      # integer altitude against a BOOL (reference quirk)
            bonus = 0.0 if hit else self.healthy
            if self.task == "hovering_control":
                bonus -= 1.0 * np.linalg.norm(st.v) + 1.0 * np.linalg.norm(st.w)
                z_move = abs(self.z0 - st.p[2])
                bonus += 10 if z_move < 0.5 else max(-20, 0.5 - z_move)
            reward += bonus
            if hit:
                done = True
                self.ct = 0
        if self.ct == self.nt:
            done = True
            self.ct = 0
        return obs, reward, done, {}


def _worker(args):
    """One process = one env stepping `seconds` of wall clock with U(0.1, 15) actions; returns env-steps done."""
    import time
    task, dt, nt, seed, seconds, warm = args
    np.random.seed(1000 + seed)
    env = NumpyQuadrotorEnv(dt=dt, nt=nt, seed=seed, task=task)
    env.reset()
    rng = np.random.RandomState(seed)
    n = 0
    tw = time.perf_counter()
    while time.perf_counter() - tw < warm:
        env.step(rng.uniform(0.1, 15.0, 4).astype(np.float32))
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        try:
            _, _, done, _ = env.step(rng.uniform(0.1, 15.0, 4).astype(np.float32))
        except Exception:
            done = True
        if done:
            env.reset()
        n += 1
    return n, time.perf_counter() - t0


def measure_throughput_detail(task="velocity_control", dt=0.005, nt=1000, seconds=5.0, processes=None,
                              warmup_seconds=0.2):
    """One env per process (multiprocessing, spawn), each stepping `seconds` of wall clock after `warmup_seconds`.
    -> {"value": aggregate env-steps/s, "rates": per-process env-steps/s, "wall_s": slowest process's timed wall}."""
    import multiprocessing as mp
    import os
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(v, "1")        # np.linalg.inv on a 3x3 must not wake a BLAS thread pool per process
    cores = processes or len(os.sched_getaffinity(0))
    # short table (nt_eff) so that process start-up is not dominated by the velocity-table generation
    nt_eff = min(nt, 50)
    ctx = mp.get_context("spawn")   # the caller may hold a CUDA context: never fork it
    with ctx.Pool(cores) as pool:
        res = pool.map(_worker, [(task, dt, nt_eff, k, seconds, warmup_seconds) for k in range(cores)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {"value": total / wall, "rates": [r[0] / r[1] for r in res], "wall_s": wall, "processes": cores}


def measure_throughput(task="velocity_control", dt=0.005, nt=1000, seconds=5.0, processes=None, warmup_seconds=0.2):
    """env-steps/s of the numpy port with one env per host core (multiprocessing). -> (steps_per_s, cores)."""
    d = measure_throughput_detail(task, dt, nt, seconds, processes, warmup_seconds)
    return d["value"], d["processes"]
