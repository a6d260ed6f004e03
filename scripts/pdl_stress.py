"""Stress of the launch-to-launch tile hand-off: long chains of step kernels (graph and eager), interleaved resets /
rollouts / host steps, compared bit for bit with a fully serialised handle (MGB_PDL=0 semantics via a second process is
not needed: the fused rollout kernel is the independent reference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metagym_b200 import BatchedQuadrotor

torch.manual_seed(0)
for N in (65536, 1000, 300000):
    kw = dict(task="velocity_control", dt=0.005, nt=37, seed=list(range(8)), num_envs=N, squeeze=False, auto_reset=True,
              rng_seed=3)
    a, b = BatchedQuadrotor(**kw), BatchedQuadrotor(**kw)
    a.reset(); b.reset()
    T = 96
    acts = torch.rand((T, N, 4), device="cuda") * 14.9 + 0.1
    out = b.rollout(T, actions=acts)          # reference: one fused launch
    obs = torch.empty((T, N, 19), device="cuda"); rew = torch.empty((T, N), device="cuda")
    done = torch.empty((T, N), dtype=torch.uint8, device="cuda")
    # eager chain, distinct output slots
    for t in range(T // 2):
        a.step(acts[t], out=(obs[t], rew[t], done[t]))
    # graph chain for the second half
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for t in range(T // 2, T):
                a.step(acts[t], out=(obs[t], rew[t], done[t]))
    g.replay(); torch.cuda.synchronize()
    ok = torch.equal(obs, out["obs"]) and torch.equal(rew, out["rew"]) and torch.equal(done, out["done"])
    sa, sb = a.state_dict(), b.state_dict()
    ok = ok and torch.equal(sa["state"], sb["state"]) and torch.equal(sa["ct"], sb["ct"])
    # same output tensor reused by consecutive steps: the last step must win
    for rep in range(50):
        a.reset(); b.reset()
        for t in range(8):
            o, r, d, _ = a.step(acts[t])
        ref = b.rollout(8, actions=acts[:8])
        ok = ok and torch.equal(o, ref["obs"][7]) and torch.equal(r, ref["rew"][7])
    # graph replayed many times back to back (tickets keep counting across replays)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    print("N=%d tile hand-off chain == fused rollout: %s" % (N, ok))
    assert ok
    a.close(); b.close()
print("pdl stress ok")
