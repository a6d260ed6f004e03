#!/usr/bin/env python
"""bench.py -- env-steps/sec of the quadrotor hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the config the metric is quoted on): Quadrotor velocity_control, 65 536 envs per
GPU, dt = 0.005 = 5 Euler substeps of 1 ms (the reference's only integrator), nt = 1000, 64 velocity
tasks (seeds 0..63), U(0.1, 15) random actions, auto-reset on.  One "step" = one env.step() of every env of the batch =
ONE launch of the state-update kernel through the C ABI (mgb_quad_step).

  value     device-resident: K steps replayed from CUDA graphs of 32 steps each; actions are read from and observations
            written to rollout buffers [32, n, .] (obs 160 MB > 126 MB L2, so inputs/outputs are never L2-hot; the 6 MB
            recurrent state stays in L2 between consecutive steps, which is the nature of the workload).  Timed with
            CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  roofline  algorithmic bytes (281 B/env-step) x envs / average launch duration, against the measured
            copy bandwidth in MEASURED_PEAKS.json.  `streaming` repeats it with 4 194 304 envs (state streams from HBM).
  e2e       the same step through mgb_quad_step_host with PINNED HOST buffers: H2D of the actions, kernel, D2H of
            obs/reward/done inside the timed region (wall clock, synchronous call).
  cpu_baseline  the single-env numpy port of the reference loop (oracle/quadrotor_np.py, bit-identical to the
            reference on the golden vectors), one env per host core, ~10 s.
  --impl reference  times that same numpy port with all host cores (the reference is pure Python and does not exist on
            the GPU box; oracle/_ref cannot be built because there is nothing to compile).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (quadrotor 6-DoF, 65k envs)"
N_ENVS_PER_GPU = 65536
DT, NT, N_TASKS = 0.005, 1000, 64
TASK = "velocity_control"
FLOPS_PER_STEP = 350 * 5 + 60   # dt 0.005 = 5 substeps
BYTES_PER_STEP = 281          # read state 88 + ct 4 + action 16; write state 88 + ct 4 + obs 76 + rew 4 + done 1
GRAPH_STEPS = 32              # steps per CUDA graph = slots of the rollout buffer
# dram__bytes_read.sum + dram__bytes_write.sum of ONE quad_step_kernel launch at this shape, from the committed
# `ncu --set full` capture (not measured by this script; ncu flushes caches, so the L2-resident state is read from DRAM
# once and the 12.5 MB of outputs had not been written back when the kernel ended).  Algorithmic: 65536 x 281 = 18.4 MB.
NCU_TRAFFIC_BYTES = 7680768
NCU_TRAFFIC_SOURCE = "profiles/r1_ncu_quad_step_wide_65k.txt (4M-env launch: 1186.9 MB vs 1178.6 MB algorithmic, r1_ncu_quad_step_4m.txt)"
WORKLOAD = ("quadrotor velocity_control, %d envs/GPU, dt=0.005 (5 Euler substeps of 1 ms), nt=1000, "
            "64 velocity tasks, U(0.1,15) actions, auto-reset" % N_ENVS_PER_GPU)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 100 ms while the load runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append((time.time(), parts))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.15)]
        if not rows:
            rows = [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    return len(os.sched_getaffinity(0))


def cpu_baseline(seconds):
    from oracle.quadrotor_np import measure_throughput
    v, cores = measure_throughput(task=TASK, dt=DT, nt=NT, seconds=seconds)
    return {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "numpy single-env port of the reference loop (oracle/quadrotor_np.py), one env per core, "
                      "%.0f s wall per process, velocity_control dt=0.005, U(0.1,15) actions" % seconds}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    # K "steps" = K equal slices of ONE bounded run (40 s timed after 5 s warm-up), so any K/W ends within a minute
    timed_s, warm_s = 40.0, 5.0
    per_step_s = timed_s / max(1, args.steps)
    from oracle.quadrotor_np import measure_throughput
    value, cores = measure_throughput(task=TASK, dt=DT, nt=NT, seconds=timed_s, warmup_seconds=warm_s)
    vals = [value]
    n_local = N_ENVS_PER_GPU
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_local * args.gpus / value * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm: the K steps are K slices (%.3f s each) of one bounded 40 s "
                                                 "run after 5 s warm-up, one numpy env per host core; ms_per_step is "
                                                 "the time this arm would need for one 65 536-env batch step" % per_step_s},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "numpy single-env port (bit-identical to the reference on the golden vectors), "
                                   "%d processes x %.0f s" % (cores, timed_s)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def maze_extras(torch, dev, peak):
    """env-steps/s of MetaMazeDiscrete3D (15x15, 128x128, uint8, 1024 envs = config 4 per GPU) and MetaMaze2D."""
    import numpy as np
    from metagym_b200 import BatchedMetaMaze2D, BatchedMetaMazeDiscrete3D, MazeTaskSampler
    rs = np.random.RandomState(0)
    tasks = [MazeTaskSampler(n=15, allow_loops=True, crowd_ratio=0.35, rng=rs) for _ in range(64)]
    out = {}
    for name, make, n, nbytes in (
            ("maze3d_15x15_128x128_u8_1024envs",
             lambda: BatchedMetaMazeDiscrete3D(resolution=(128, 128), max_steps=200, task_type="SURVIVAL", num_envs=1024,
                                               device=dev.index, squeeze=False, auto_reset=True, obs_dtype="uint8"),
             1024, 128 * 128 * 3 + 1630),
            ("maze2d_15x15_g1_16384envs",
             lambda: BatchedMetaMaze2D(max_steps=200, task_type="ESCAPE", view_grid=1, num_envs=16384, device=dev.index,
                                       squeeze=False, auto_reset=True), 16384, 160)):
        env = make()
        env.set_task(tasks)
        env.reset()
        acts = torch.randint(0, 4, (16, n), device=dev, dtype=torch.int32)
        for t in range(5):
            env.step(acts[t % 16])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 50
        e0.record()
        for t in range(iters):
            env.step(acts[t % 16])
        e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / iters
        out[name] = {"value": n / us * 1e6, "unit": "env-steps/s", "us_per_step": us,
                     "algorithmic_bytes_per_env_step": nbytes, "frac_of_measured_hbm": n * nbytes / us * 1e-3 / peak,
                     "parity": "bit-exact vs reference golden episodes (tests/test_maze_gpu.py)"}
        if hasattr(env, "rollout"):          # T steps per launch, device-drawn uniform actions
            T = 32 if nbytes < 1000 else 16
            bufs = env.rollout(T, act_seed=3, want_actions=True)
            for _ in range(2):
                env.rollout(T, act_seed=3, out=bufs)
            e0.record()
            for _ in range(8):
                env.rollout(T, act_seed=3, out=bufs)
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / (8 * T)
            out[name + "_fused_rollout"] = {"value": n / us * 1e6, "unit": "env-steps/s", "us_per_step": us, "T": T,
                                            "frac_of_measured_hbm": n * nbytes / us * 1e-3 / peak,
                                            "parity": "bit-exact vs T single steps (tests/test_maze_gpu.py)"}
        env.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4096)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS_PER_GPU, help="envs per GPU (default: the BASELINE config)")
    ap.add_argument("--no-extras", action="store_true", help="skip streaming / fused / e2e / cpu legs")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert args.warmup >= 3, "W >= 3 warm-up steps are required"
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from metagym_b200 import BatchedQuadrotor, _lib
    from metagym_b200.rollout import all_gather_rollout, rollout_bytes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: metagym_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    base = rank * n
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    env = BatchedQuadrotor(task=TASK, dt=DT, nt=NT, seed=list(range(N_TASKS)), num_envs=n, device=local_rank,
                           squeeze=False, auto_reset=True, rng_seed=0, env_index_base=base)
    env.reset()
    D = env.obs_dim
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    G = GRAPH_STEPS
    acts = torch.rand((G, n, 4), device=dev, generator=gen) * 14.9 + 0.1
    obs = torch.empty((G, n, D), dtype=torch.float32, device=dev)
    rew = torch.empty((G, n), dtype=torch.float32, device=dev)
    done = torch.empty((G, n), dtype=torch.uint8, device=dev)

    def enqueue(t):
        env.step(acts[t % G], out=(obs[t % G], rew[t % G], done[t % G]))

    stream = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for t in range(max(3, min(W, G))):
            enqueue(t)
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for t in range(G):
                enqueue(t)
    torch.cuda.synchronize(dev)

    def run_steps(k):
        """k steps on the current stream: whole graphs of G steps, remainder as single launches."""
        for _ in range(k // G):
            graph.replay()
        for t in range(k % G):
            enqueue(t)

    run_steps(W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = n * world * K / (ms * 1e-3)
    us_per_launch = ms * 1e3 / K
    peak, peak_src = measured_peak_gbs()
    achieved = n * BYTES_PER_STEP / (us_per_launch * 1e-6) / 1e9

    # keep the same load running long enough for nvidia-smi to see it (the timed region itself can be milliseconds)
    t_load0 = time.time()
    while time.time() - t_load0 < 1.5:
        run_steps(G * 64)
        torch.cuda.synchronize(dev)
    t_load1 = time.time()
    clocks = sampler.stop(t_wall0, t_load1) if rank == 0 else None
    finite = bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())

    extras = {}
    e2e = None
    if not args.no_extras:
        # ---- end to end through the host-buffer C-ABI entry point, pinned buffers, copies inside the timed region
        h_act = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        h_act.copy_(acts[0].cpu())
        h_obs = torch.empty((n, D), dtype=torch.float32).pin_memory()
        h_rew = torch.empty((n,), dtype=torch.float32).pin_memory()
        h_done = torch.empty((n,), dtype=torch.uint8).pin_memory()
        for _ in range(5):
            env.step_host_buffers(h_act, h_obs, h_rew, h_done)
        Ke = max(20, min(K, 200))
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            env.step_host_buffers(h_act, h_obs, h_rew, h_done)
        torch.cuda.synchronize(dev)
        dt_e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n * world * Ke / dt_e, "unit": "env-steps/s", "h2d_bytes_per_step": n * 16,
               "d2h_bytes_per_step": n * (D * 4 + 4 + 1), "steps": Ke, "timer": "host wall clock around the synchronous "
               "mgb_quad_step_host calls (pinned host buffers), max over ranks",
               "result_checksum": float(h_rew.double().sum())}
        assert bool(torch.isfinite(h_obs).all())

        # ---- fused T-step rollout kernel (state in registers), same buffers
        out = {"obs": obs, "rew": rew, "done": done, "act": None}
        for _ in range(3):
            env.rollout(G, actions=acts, out=out)
        barrier()
        reps = max(4, K // G)
        e0.record()
        for _ in range(reps):
            env.rollout(G, actions=acts, out=out)
        e1.record()
        barrier()
        msf = max_over_ranks(e0.elapsed_time(e1))
        extras["fused_rollout"] = {"value": n * world * reps * G / (msf * 1e-3), "unit": "env-steps/s",
                                   "T": G, "launches": reps,
                                   "note": "mgb_quad_rollout: T steps per launch, state held in registers"}
        if world > 1:
            # ---- the one collective of the path: all-gather of a rollout chunk for the learner (NCCL over NVLink)
            chunk = {"obs": obs, "act": acts, "rew": rew, "done": done}
            for _ in range(3):
                all_gather_rollout(chunk)
            barrier()
            e0.record()
            for _ in range(5):
                gathered = all_gather_rollout(chunk)
            e1.record()
            barrier()
            msg = max_over_ranks(e0.elapsed_time(e1)) / 5
            nbytes = rollout_bytes(chunk)
            extras["rollout_allgather"] = {"ms": msg, "bytes_per_rank": nbytes,
                                           "busbw_GBps": nbytes * (world - 1) / (msg * 1e-3) / 1e9,
                                           "gathered_envs": int(gathered["obs"].shape[1])}
            # ---- the same exchange done by the rollout kernel itself: every output is also stored into the other ranks'
            # receive arenas over NVLink (PeerArena); the only cross-rank call left is a one-element rendezvous
            arenas = []
            try:
                from metagym_b200.rollout import PeerArena
                fields = {"obs": ((G, n, D), torch.float32), "act": ((G, n, 4), torch.float32),
                          "rew": ((G, n), torch.float32), "done": ((G, n), torch.uint8)}
                arenas = [PeerArena(fields, dev) for _ in range(2)]
                tick = [0]

                def chunk_peer():
                    ar = arenas[tick[0] & 1]
                    tick[0] += 1
                    ar.attach(env)
                    env.rollout(G, actions=None, act_seed=3, out=ar.views)
                    return ar.sync()

                for _ in range(3):
                    views = chunk_peer()
                barrier()
                reps = max(4, K // G)
                e0.record()
                for _ in range(reps):
                    views = chunk_peer()
                e1.record()
                barrier()
                msp = max_over_ranks(e0.elapsed_time(e1))
                own = bool(torch.equal(views["obs"][rank], arenas[(tick[0] - 1) & 1]["obs"]))
                extras["fused_rollout_peer_gather"] = {
                    "value": n * world * reps * G / (msp * 1e-3), "unit": "env-steps/s", "T": G,
                    "chunk_bytes_per_rank": arenas[0].payload_bytes(),
                    "ingress_GBps_per_gpu": arenas[0].payload_bytes() * (world - 1) * reps / (msp * 1e-3) / 1e9,
                    "gathered_shape": list(views["obs"].shape), "own_slot_ok": own,
                    "note": "quad_rollout_kernel<.,1>: outputs stored into every rank's arena by the kernel (NVLink "
                            "peer stores), rendezvous = 1-element all-reduce; no NCCL on the data"}
            except Exception as ex:            # reported, never fatal for the headline numbers
                extras["fused_rollout_peer_gather"] = {"error": repr(ex)[:300]}
            finally:
                env.set_mirrors([])
                for ar in arenas:
                    try:
                        ar.close()
                    except Exception:
                        pass
        env.close()
        del env, obs, rew, done, acts
        torch.cuda.empty_cache()
        # ---- streaming variant: 4 194 304 envs, the state (403 MB) no longer fits L2 and streams from HBM
        if rank == 0:
            ns = 4194304
            big = BatchedQuadrotor(task=TASK, dt=DT, nt=NT, seed=list(range(N_TASKS)), num_envs=ns, device=local_rank,
                                   squeeze=False, auto_reset=True)
            big.reset()
            a2 = torch.rand((2, ns, 4), device=dev, generator=gen) * 14.9 + 0.1
            for t in range(4):
                big.step(a2[t % 2])
            torch.cuda.synchronize(dev)
            e0.record()
            for t in range(20):
                big.step(a2[t % 2])
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / 20
            extras["streaming"] = {"envs": ns, "us_per_launch": us, "value": ns / us * 1e6, "unit": "env-steps/s",
                                   "achieved_GBps": ns * BYTES_PER_STEP / us * 1e-3,
                                   "frac": ns * BYTES_PER_STEP / us * 1e-3 / peak}
            big.close()
            del big, a2
            torch.cuda.empty_cache()
            # ---- the other half of the hot path: MetaMaze (BASELINE configs[3]/[4] per-GPU shapes), bit-exact renderer
            try:
                extras["metamaze"] = maze_extras(torch, dev, peak)
            except Exception as ex:      # the contract metric above must not depend on this leg
                extras["metamaze"] = {"error": repr(ex)[:200]}
    else:
        env.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "global_envs": n * world, "parallelism": "dp%d" % world,
                       "launch": "CUDA graphs of %d mgb_quad_step launches (programmatic dependent launch)" % G,
                       "l2": "rollout buffers obs [32,n,19] f32 = 160 MB > 126 MB L2 (inputs/outputs never L2-hot); "
                             "the 6 MB recurrent state is L2-resident by nature of the workload; see extras.streaming "
                             "for the 4M-env run whose state streams from HBM",
                       "integrator": "semi-implicit Euler substeps (the reference's integrator); BASELINE's 'RK4' has "
                                     "no reference counterpart (SURVEY.md fact 2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "kernel": "quad_step_wide_kernel<true>",
                         "bytes_per_env_step": BYTES_PER_STEP, "envs_per_launch": n,
                         "us_per_launch": us_per_launch,
                         # secondary figure SURVEY.md 8d asks for: ~350 flop per substep + ~60 per step (hand count)
                         "flops_per_env_step": FLOPS_PER_STEP,
                         "achieved_fp32_tflops": n / (us_per_launch * 1e-6) * FLOPS_PER_STEP / 1e12,
                         "fp32_peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12},
            "gpu_launches": K,
            "clocks": clocks,
            "finite_outputs": finite,
        }
        if e2e is not None:
            line["e2e"] = e2e
        line.update(extras)
        if not args.no_extras:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
