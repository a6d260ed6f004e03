"""Development aid (test infrastructure, not collected by pytest): per-step max error of the GPU quadrotor against the CPU oracle (mix) and the f64 arbiter."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from metagym_b200 import BatchedQuadrotor
from oracle import quad_oracle as qo
from util import OBS_GROUPS, STATE_GROUPS, group_rel_err

cfg = qo.make_cfg()
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quadrotor_golden.npz"))
# KAT 200
env = BatchedQuadrotor(task="velocity_control", nt=1000, seed=0, num_envs=1, squeeze=False)
act = torch.full((1, 4), 5.0, device="cuda")
ref = g["kat2_states"]
for t in range(200):
    env.step(act)
    if t % 20 == 19:
        st = env.state_dict()["state"].cpu().numpy().astype(np.float64)
        print("kat200 t=%d err=%.2e  w=%.6f ref_w=%.6f vz=%.7f ref=%.7f" % (t, group_rel_err(st, ref[t][None], STATE_GROUPS), st[0, 9], ref[t][9], st[0, 5], ref[t][5]))
env.close()
for task, dt in (("hovering_control", 0.01), ("velocity_control", 0.005)):
    n, nt = 4096, 30
    rng = np.random.RandomState(1234 + n)
    env = BatchedQuadrotor(task=task, dt=dt, nt=nt, seed=[0, 1, 2], num_envs=n, squeeze=False)
    noise = rng.random_sample((n, 12))
    env.reset(noise=noise)
    sm, s64 = qo.reset_state(None, noise), qo.reset_state(None, noise)
    cm, c64 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    kw = {}
    if task == "velocity_control":
        kw = dict(targets=env.velocity_targets.cpu().numpy(), env2task=env.env2task.cpu().numpy())
    for t in range(20):
        act = rng.uniform(-1.0, 16.0, (n, 4)).astype(np.float32)
        obs, rew, done, _ = env.step(torch.as_tensor(act).cuda())
        om, *_ = qo.env_step(cfg, sm, cm, act, task, dt, nt, mode="mix", **kw)
        o64, *_ = qo.env_step(cfg, s64, c64, act, task, dt, nt, mode="f64", **kw)
        o = obs.cpu().numpy()
        print("%s t=%d gpu-mix=%.2e gpu-f64=%.2e mix-f64=%.2e" % (task, t, group_rel_err(o[:, :16], om[:, :16], OBS_GROUPS),
              group_rel_err(o[:, :16], o64[:, :16], OBS_GROUPS), group_rel_err(om[:, :16], o64[:, :16], OBS_GROUPS)))
    env.close()
