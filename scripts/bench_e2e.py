"""Host-buffer (e2e) step rate of mgb_quad_step_host for the MGB_HOST_ZEROCOPY modes, one subprocess per mode.
    python scripts/bench_e2e.py            # modes 1 (zero-copy), 2 (hybrid: DMA actions, zero-copy outputs), 0 (copies)
"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from metagym_b200 import BatchedQuadrotor
    n = 65536
    env = BatchedQuadrotor(task="velocity_control", dt=0.005, nt=1000, seed=list(range(64)), num_envs=n, device=0,
                           squeeze=False, auto_reset=True)
    env.reset()
    D = env.obs_dim
    h_act = (torch.rand((n, 4)) * 14.9 + 0.1).pin_memory()
    h_obs = torch.empty((n, D)).pin_memory()
    h_rew = torch.empty((n,)).pin_memory()
    h_done = torch.empty((n,), dtype=torch.uint8).pin_memory()
    for _ in range(20):
        env.step_host_buffers(h_act, h_obs, h_rew, h_done)
    torch.cuda.synchronize()
    K = 300
    t0 = time.perf_counter()
    for _ in range(K):
        env.step_host_buffers(h_act, h_obs, h_rew, h_done)
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": os.environ.get("MGB_HOST_ZEROCOPY", "1"), "us_per_step": dt / K * 1e6,
                      "env_steps_per_s": n * K / dt, "checksum": float(h_rew.double().sum())}), flush=True)
    sys.exit(0)
for mode in ("1", "2", "0", "2", "1"):
    env = dict(os.environ, MGB_HOST_ZEROCOPY=mode)
    r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
