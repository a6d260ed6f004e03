"""Device-side throughput of the MetaMaze kernels at the BASELINE config-4 / config-5 per-GPU shapes (development +
profiles/; bench.py carries the contract metric).  Prints one JSON line per case."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from metagym_b200 import BatchedMetaMaze2D, BatchedMetaMazeDiscrete3D, MazeTaskSampler


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    rs = np.random.RandomState(0)
    tasks = [MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs) for _ in range(64)]
    cases = [("maze3d_u8_1024", "3D", 1024, "uint8"), ("maze3d_i32_1024", "3D", 1024, "int32"),
             ("maze3d_u8_8192", "3D", 8192, "uint8"), ("maze2d_16384", "2D", 16384, None),
             ("maze2d_1048576", "2D", 1048576, None)]
    for name, kind, n, dt in cases:
        if args.only and args.only not in name:
            continue
        if kind == "3D":
            env = BatchedMetaMazeDiscrete3D(resolution=(128, 128), max_steps=200, task_type="SURVIVAL", num_envs=n,
                                            squeeze=False, auto_reset=True, obs_dtype=dt)
            bytes_per_step = 128 * 128 * 3 * (1 if dt == "uint8" else 4) + 1600 + 30
        else:
            env = BatchedMetaMaze2D(max_steps=200, task_type="ESCAPE", view_grid=1, num_envs=n, squeeze=False,
                                    auto_reset=True)
            bytes_per_step = 160
        env.set_task(tasks)
        env.reset()
        acts = torch.randint(0, 4, (16, n), device="cuda", dtype=torch.int32)
        k = [0]

        def step():
            env.step(acts[k[0] % 16])
            k[0] += 1
        for _ in range(3):
            step()
        us = timeit(step, args.iters)
        gbps = n * bytes_per_step / us * 1e-3
        print(json.dumps({"case": name, "envs": n, "us_per_step": us, "env_steps_per_s": n / us * 1e6,
                          "algorithmic_bytes_per_env_step": bytes_per_step, "achieved_GBps": gbps,
                          "frac_of_measured_hbm": gbps / peak()}), flush=True)
        if kind == "3D" and dt == "uint8":
            T = 16 if n <= 1024 else 4

            def roll3():
                env.rollout(T, act_seed=3, out=bufs3)
            bufs3 = env.rollout(T, act_seed=3, want_actions=True)
            roll3()
            us_r = timeit(roll3, 4) / T
            print(json.dumps({"case": name + "_fused_rollout", "envs": n, "T": T, "us_per_step": us_r,
                              "env_steps_per_s": n / us_r * 1e6, "achieved_GBps": n * bytes_per_step / us_r * 1e-3,
                              "frac_of_measured_hbm": n * bytes_per_step / us_r * 1e-3 / peak()}), flush=True)
            del bufs3
        if kind == "2D":
            T = 32 if n <= 65536 else 8

            def roll():
                env.rollout(T, act_seed=3, out=bufs)
            bufs = env.rollout(T, act_seed=3, want_actions=True)
            for _ in range(2):
                roll()
            us_r = timeit(roll, max(4, args.iters // 2)) / T
            print(json.dumps({"case": name + "_fused_rollout", "envs": n, "T": T, "us_per_step": us_r,
                              "env_steps_per_s": n / us_r * 1e6, "achieved_GBps": n * bytes_per_step / us_r * 1e-3,
                              "frac_of_measured_hbm": n * bytes_per_step / us_r * 1e-3 / peak()}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
