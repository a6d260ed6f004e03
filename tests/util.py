"""Shared comparison helpers for the parity tests."""
import numpy as np

# state row = p3 v3 w3 prop4 R9
STATE_GROUPS = [(0, 3), (3, 6), (6, 9), (9, 13), (13, 22)]
# obs row = b_v3 b_p3 acc3 gyro3 (pitch roll yaw) z [target3]
OBS_GROUPS = [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 16)]
TASK_NAMES = {0: "no_collision", 1: "hovering_control", 2: "velocity_control"}


def group_rel_err(a, ref, groups, floor=1e-3):
    """max over entries of |a-ref| / max(group inf-norm of ref, floor): the '1e-5 relative fp32' metric of the
    north star, taken per physical vector (a 1e-15 m position component is compared against the vector's size)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    a = a.reshape(-1, a.shape[-1])
    ref = ref.reshape(-1, ref.shape[-1])
    worst = 0.0
    for lo, hi in groups:
        scale = np.maximum(np.abs(ref[:, lo:hi]).max(axis=1, keepdims=True), floor)
        err = np.abs(a[:, lo:hi] - ref[:, lo:hi]) / scale
        if err.size:
            worst = max(worst, float(np.nanmax(err)))
        assert not np.isnan(a[:, lo:hi]).any()
    return worst


def scalar_rel_err(a, ref, floor=1.0):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float((np.abs(a - ref) / np.maximum(np.abs(ref), floor)).max()) if a.size else 0.0


def golden_run(g, name):
    """dict view of one recorded run of tests/golden/quadrotor_golden.npz."""
    pre = name + "."
    d = {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
    task, dt, nt, seed = d["meta"]
    d["task"], d["dt"], d["nt"], d["seed"] = TASK_NAMES[int(task)], float(dt), int(nt), int(seed)
    return d


QUAD_RUNS = ["hover_a", "hover_b", "hover_fall", "nocol_fall", "nocol_a", "vel_a", "vel_b", "vel_c"]
