"""Measure (on a B200, through gpurun) how fast the GPU quadrotor drifts from the reference's golden free-running
episodes, and write the running-max curve tests/test_quadrotor_gpu.py asserts against (x2).

    python tests/golden/measure_free_run_envelope.py          # -> gpurun_out/free_run_envelope.json
    cp gpurun_out/free_run_envelope.json tests/golden/free_run_envelope.json

Error metric = tests/util.py group_rel_err over the 16 sensor observations (per physical vector, angles against 1 rad,
z against its 5 m offset).  curve[j] = max over the recorded episodes of the error at step j of an episode, made
monotone (running max): a free run can only be asserted as tightly as its worst earlier step.
Also records the same curve for the CPU oracle in float64 (the reference's own float32 noise: SURVEY.md 8c).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch
from metagym_b200 import BatchedQuadrotor
from util import OBS_GROUPS, golden_run, group_rel_err

g = np.load(os.path.join(HERE, "quadrotor_golden.npz"))
RUNS = ["hover_a", "hover_fall", "nocol_a", "vel_a", "vel_c"]
curve = np.zeros(0)
per_run = {}
for name in RUNS:
    r = golden_run(g, name)
    kw = dict(dt=r["dt"], nt=r["nt"])
    if r["task"] == "velocity_control":
        kw["seed"] = r["seed"]
    env = BatchedQuadrotor(task=r["task"], num_envs=1, device=0, squeeze=False, **kw)
    ep = r["ep"]
    errs = []
    for k in range(int(ep.max()) + 1):
        idx = np.nonzero(ep == k)[0]
        env.reset(noise=r["reset_noise"][k][None])
        for j, i in enumerate(idx):
            obs, rew, done, _ = env.step(torch.as_tensor(r["act"][i][None]).cuda())
            e = group_rel_err(obs.cpu().numpy()[:, :16], r["obs"][i][None, :16], OBS_GROUPS)
            if j >= len(errs):
                errs.append(e)
            else:
                errs[j] = max(errs[j], e)
    env.close()
    per_run[name] = errs
    if len(errs) > len(curve):
        curve = np.concatenate([curve, np.zeros(len(errs) - len(curve))])
    curve[:len(errs)] = np.maximum(curve[:len(errs)], errs)
mono = np.maximum.accumulate(curve)
# ---- the same for test_random_batch_vs_oracle: 4096 envs, 20 free steps, actions U(-1, 16), vs the CPU oracle (mix mode)
from oracle import quad_oracle as qo
cfg = qo.make_cfg()
batch = np.zeros(20)
for task, dt in (("hovering_control", 0.01), ("velocity_control", 0.005), ("no_collision", 0.01)):
    for n in (1, 127, 129, 4096):
        rng = np.random.RandomState(1234 + n)
        nt = 30
        env = BatchedQuadrotor(task=task, num_envs=n, device=0, squeeze=False, dt=dt, nt=nt, seed=[0, 1, 2])
        noise = rng.random_sample((n, 12))
        env.reset(noise=noise)
        state = qo.reset_state(None, noise)
        ct = np.zeros(n, np.int32)
        kw = {}
        if task == "velocity_control":
            kw = dict(targets=env.velocity_targets.cpu().numpy(), env2task=env.env2task.cpu().numpy())
        for t in range(20):
            act = rng.uniform(-1.0, 16.0, (n, 4)).astype(np.float32)
            obs, rew, done, _ = env.step(torch.as_tensor(act).cuda())
            o_ref = qo.env_step(cfg, state, ct, act, task, dt, nt, mode="mix", **kw)[0]
            batch[t] = max(batch[t], group_rel_err(obs.cpu().numpy()[:, :16], o_ref[:, :16], OBS_GROUPS))
        env.close()
batch = np.maximum.accumulate(batch)
out = {"random_batch_vs_oracle_running_max": [float(x) for x in batch],"metric": "group_rel_err over obs[:16] (tests/util.py), GPU free run vs reference golden episodes",
       "runs": RUNS, "curve_running_max": [float(x) for x in mono], "per_run_max": {k: float(max(v)) for k, v in per_run.items()},
       "per_run_len": {k: len(v) for k, v in per_run.items()}, "device": torch.cuda.get_device_name(0)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "free_run_envelope.json"), "w") as f:
    json.dump(out, f)
print("steps", len(mono), "err@1,10,50,100,end:", [float(mono[min(i, len(mono) - 1)]) for i in (0, 9, 49, 99, len(mono) - 1)])
print(out["per_run_max"])
print("random batch vs oracle, err@1,5,10,20:", [float(batch[i]) for i in (0, 4, 9, 19)])
