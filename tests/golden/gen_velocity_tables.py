"""Generate tests/golden/velocity_tables_golden.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/gen_velocity_tables.py

define_velocity_control_task(dt=0.005, nt=1000, seed) (metagym/quadrotor/quadrotorsim.py:306-319) for seeds 0..3 -- the
table shape BASELINE.json's configs[2] (and bench.py) uses: 1000 free-running steps of 5 substeps from the zero state
with numpy-seeded U(0.1, 15) actions, global_velocity recorded after every step.  These rows feed obs[16:19] and the
velocity_control reward, so the GPU generator (mgb_quad_make_targets) is pinned on them at the benchmarked size, not
only at nt=40.  Also records how fast the float32 reference and its float64 restatement drift apart along these
trajectories (the measured envelope the GPU test asserts, x2).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _refload  # noqa: E402


def main():
    ns = _refload.load_reference()
    out = {}
    seeds = [0, 1, 2, 3]
    tabs = []
    for s in seeds:
        sim = ns.QuadrotorSim()                           # Quadrotor.__init__: get_config -> define task (env.py:62-69)
        sim.get_config(ns.quad_config)
        tab = np.asarray(sim.define_velocity_control_task(0.005, 1000, s), dtype=np.float32)
        assert tab.shape == (1000, 3)
        tabs.append(tab)
    out["seeds"] = np.asarray(seeds, dtype=np.int64)
    out["tables"] = np.stack(tabs)
    out["meta"] = np.asarray([0.005, 1000], dtype=np.float64)
    # f32-vs-f64 drift of the same trajectories: the oracle's float64 arbiter driven by the same action draws
    from oracle import quad_oracle as qo
    cfg = qo.make_cfg()
    env = []
    for s in seeds:
        rs = np.random.RandomState(s)
        acts = rs.uniform(low=0.1, high=15.0, size=(1000, 4)).astype(np.float32)
        st = qo.zero_state(1)
        tab64 = np.zeros((1000, 3))
        for t in range(1000):
            qo.sim_step(cfg, st, acts[t][None], 5, "f64")
            tab64[t] = st[0, 3:6]
        env.append(np.abs(tab64 - out["tables"][len(env)]).max(axis=1) / np.maximum(np.abs(tab64).max(axis=1), 1.0))
    out["f32_vs_f64_envelope"] = np.stack(env)            # [seed, t]: the reference's own float32 noise along the table
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "velocity_tables_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})
    print("f32-vs-f64 envelope at t=10,100,500,999:", out["f32_vs_f64_envelope"].max(axis=0)[[9, 99, 499, 999]])


if __name__ == "__main__":
    main()
