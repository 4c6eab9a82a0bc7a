#!/usr/bin/env python
"""bench.py -- DQN train-steps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload single|population|replay]

Default workload = BASELINE.json configs[1]: single-agent fused dueling double-DQN train step,
synthetic 1M-transition replay, batch 64, D=8, A=4, hidden (32,64), AdamW(2e-4), gamma .99 on one B200.
A "step" is one train step (one minibatch: sample -> targets -> loss -> backward -> Adam).  The single
agent does not shard ("replicas only"), so --gpus N runs N independent replicas, one process
per GPU, no data-path collective; `value` = all ranks' steps / max-over-ranks device time.

  value  : steps/s with everything resident in HBM, K steps fused per persistent launch
  e2e    : steps/s through the reference-facing API (ReplayBuffer.add x train_frequency from host
           memory, Agent._step(), loss read back) -- host<->device copies inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md "Measurement"

`--impl reference` times the CPU restatement of the reference path (oracle/, numpy; the reference's
jax/haiku/optax runtime is not installable) on the host cores -- the one place besides cpu_baseline
where bench.py executes oracle code, and only as the thing the GPU is compared against.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, A, B, N_RING = 8, 4, 64, 1_000_000
GAMMA, LR = 0.99, 2e-4
TRAIN_FREQUENCY = 4                      # Test/lunar_lander.py:30 -- env transitions stored per train step
REC_BYTES_ALGO = 2 * 4 * D + 8 + 4 + 1   # 77 B per sampled transition in the reference's dtypes
FLOP_PER_SAMPLE = 25728                  # D=8, H=(32,64): 3 forwards + backward
STEPS_PER_LAUNCH = int(os.environ.get("DQN_BENCH_KPL", "500"))   # train steps fused into one persistent launch


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled during the timed region (pynvml thread, 10 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(1.0)

    def summary(self):
        if not self.samples:
            try:   # fall back to one nvidia-smi sample
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0].split(",")
                return {"sm_mhz": int(out[0]), "sm_max_mhz": int(out[1]), "reasons": [], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synthetic(rng, n):
    """Synthetic transitions of SURVEY 8(d): s,s'~N(0,1); a~U{0..3}; r~2N(0,1); done~Bernoulli(0.01)."""
    s = rng.standard_normal((n, D), dtype=np.float32)
    a = rng.integers(0, A, n, dtype=np.int64)
    r = (2.0 * rng.standard_normal(n)).astype(np.float32)
    s2 = rng.standard_normal((n, D), dtype=np.float32)
    d = rng.random(n) < 0.01
    return s, a, r, s2, d


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of Agent._step on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_sec(max_steps, warmup, budget_s, ring=N_RING, seed=0):
    from threadpoolctl import threadpool_limits
    from oracle import dqn_oracle as O
    from oracle.agent_oracle import OracleAgent
    rng = np.random.default_rng(seed)
    params = O.init_params(rng, D, A)
    ora = OracleAgent(params, O.init_opt_state(params), O.OptSpec("adamw", LR), ring, D, GAMMA, B, seed=seed)
    data = synthetic(rng, ring)
    r = ora.replay                                              # bulk-fill (equivalent to `ring` add() calls)
    r.states[:], r.actions[:], r.rewards[:], r.observations[:], r.dones[:] = data
    r.counter = r.size = ring
    best = None
    ncores = os.cpu_count() or 1
    for threads in sorted({1, ncores}):
        with threadpool_limits(limits=threads):
            for _ in range(max(warmup, 3)):
                ora.step()
            t0 = time.perf_counter()
            done = 0
            while done < max_steps and time.perf_counter() - t0 < budget_s / 2:
                ora.step()
                done += 1
            dt = time.perf_counter() - t0
        rate = done / dt
        if best is None or rate > best[0]:
            best = (rate, threads, done, dt)
    return best


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    rate, threads, done, dt = cpu_reference_steps_per_sec(args.steps, args.warmup, budget_s=120.0)
    sample = f"{done} oracle train steps (B=64, D=8, 1M-slot ring) in {dt:.1f} s, BLAS threads={threads}"
    line = {
        "impl": "reference", "metric": "train_steps_per_sec", "value": rate, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": done, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "replay_samples_per_sec": rate * B,
        "note": "numpy restatement of the reference step (jax/haiku/optax not installable); host cores: %d" % (os.cpu_count() or 1),
    }
    print(json.dumps(line))


def workload_config():
    return {"workload": "configs[1]: single-agent fused dueling double-DQN train step, 1M-transition synthetic replay, batch 64",
            "obs_dim": D, "num_actions": A, "hidden": [32, 64], "batch": B, "ring_slots": N_RING, "gamma": GAMMA,
            "optimizer": "adamw(2e-4, wd 1e-4)", "steps_per_launch": STEPS_PER_LAUNCH,
            "step_kernel": os.environ.get("DQN_B200_STEP_KERNEL", "auto"),
            "multi_gpu": "replicas only (single agent does not shard)",
            "l2": "ring 96 MB < 126 MB L2: L2 flushed (512 MB write) before each timed region; every step gathers 64 random, mostly first-touch records"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def build_agent(dqn_b200, device, seed, session=False):
    rng = np.random.default_rng(seed)
    model = dqn_b200.Model(A)
    params = model.init(rng, np.zeros((1, D), np.float32))
    opt = dqn_b200.adamw(LR)
    agent = dqn_b200.Agent(network=model, params=params, optimizer=opt, opt_state=opt.init(params), env=None,
                           buffer_size=N_RING, obs_shape=(N_RING, D), ac_shape=(N_RING,), gamma=GAMMA, epsilon=1.0,
                           epsilon_decay_rate=0.99, min_epsilon=0.15, max_episodes=10000, max_steps=1500,
                           training_start=250, batch_size=B, train_frequency=TRAIN_FREQUENCY, back_up_frequency=50,
                           replace_frequency=20, reward_to_reach=230.0, num_actions=A,
                           saving_directory="/tmp/dqn_b200_bench", device=device, seed=seed, session=session)
    data = synthetic(rng, N_RING)
    for o in range(0, N_RING, 250_000):
        agent._replay_buffer.add_many(*[x[o:o + 250_000] for x in data])
    agent._engine.synchronize()
    return agent, data


def flush_l2(torch, device):
    buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    buf.fill_(1)
    torch.cuda.synchronize(device)
    del buf


def timed_fused(torch, agent, steps, device):
    """K train steps, STEPS_PER_LAUNCH per persistent launch; CUDA events on the launching stream."""
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    start.record()
    left = steps
    while left > 0:
        k = min(left, STEPS_PER_LAUNCH)
        agent._steps(k)
        left -= k
        launches += 1
    end.record()
    torch.cuda.synchronize(device)
    return start.elapsed_time(end) * 1e-3, launches


def run_single(args):
    import torch
    import dqn_b200
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    agent, data = build_agent(dqn_b200, local, seed=rank, session=not args.no_session)
    eng = agent._engine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- value: device-resident, fused ----------------------------------------------------------
    agent._steps(max(args.warmup, 3))
    flush_l2(torch, device)
    barrier()
    with ClockSampler(local) as clk:
        secs, launches = timed_fused(torch, agent, args.steps, device)
    barrier()
    clocks = clk.summary()
    tmax = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    secs_max = float(tmax.item())
    value = world * args.steps / secs_max

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_only": True, "steps_per_sec": value, "launches": launches}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: reference-facing API, host buffers, copies inside the timed region ------------------
    e2e_steps = int(min(max(args.steps // 20, 200), 5000))
    rb = agent._replay_buffer
    s, a, r, s2, d = [x[:TRAIN_FREQUENCY * (e2e_steps + 8)] for x in data]
    a_py, r_py, d_py = [int(x) for x in a], [float(x) for x in r], [bool(x) for x in d]

    def e2e_loop(n, off=0):
        last = 0.0
        for i in range(n):
            for j in range(TRAIN_FREQUENCY):                      # q_agent.py:182  one add() per env transition
                k = off + i * TRAIN_FREQUENCY + j
                rb.add(s[k], a_py[k], r_py[k], s2[k], d_py[k])
            agent._step()                                          # q_agent.py:187
            last = eng.last_loss()                                 # device -> host read of the step's loss
        return last

    e2e_loop(8)
    eng.synchronize()                       # (session mode: retire the resident kernel before other work uses the stream)
    flush_l2(torch, device)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps, off=8 * TRAIN_FREQUENCY)
    eng.synchronize()                       # inside the timed region: the session's state is back in HBM
    e1.record()
    torch.cuda.synchronize(device)
    e2e_wall = time.perf_counter() - t0
    e2e_secs = max(e0.elapsed_time(e1) * 1e-3, e2e_wall)
    te = torch.tensor([e2e_secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (dqn_train_fused_kernel), measured live ---------------
    peak, peak_src = measured_peaks()
    launch_s = secs / launches
    steps_per_launch = args.steps / launches
    algo_bytes = steps_per_launch * (B * REC_BYTES_ALGO + 4)          # gathered records + one loss store per step
    achieved = algo_bytes / launch_s / 1e9
    sm_hz = (clocks["sm_mhz"] or 1965) * 1e6
    fp32_one_sm = 128 * 2 * sm_hz / 1e9
    flops = B * FLOP_PER_SAMPLE / (secs / args.steps) / 1e9
    cluster = args.step_kernel in ("auto", "cluster")
    n_sm = 4 if cluster else 1
    kname = "dqn_train_cluster_kernel<4>" if cluster else "dqn_train_fused_kernel<4>"
    traffic = None
    try:      # DRAM bytes of one launch from the committed ncu --set full capture, scaled to this run's steps per launch
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f)[kname]
        traffic = tr["dram_bytes_per_launch"] / tr["steps_per_launch"] * steps_per_launch
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kname, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": launch_s * 1e3,
                "note": "latency-bound by construction: one agent's steps are a serial chain (step t+1 needs theta_t) on %s; "
                        "theta/theta^-/grads stay in shared memory, so the only HBM traffic is 64 gathered records per step "
                        "(prefetched one step ahead)" % ("a 4-CTA cluster (4 of 148 SMs)" if cluster else "one CTA (1 of 148 SMs)"),
                "fp32": {"achieved_gflops": flops, "sms_used": n_sm, "ffma_peak_gflops_of_sms_used": n_sm * fp32_one_sm,
                         "frac_of_sms_used": flops / (n_sm * fp32_one_sm), "flop_per_step": B * FLOP_PER_SAMPLE}}

    # ---- K=1 launches (launch-bound variant) and replay-gather throughput, for the record -----------
    extras = {}
    try:
        k1 = 2000
        eng.set_session(0)                                         # K = 1 launches: one kernel launch per Agent._step()
        agent._steps(1)
        torch.cuda.synchronize(device)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(k1):
            agent._step()
        s1.record()
        torch.cuda.synchronize(device)
        extras["k1_steps_per_sec"] = k1 / (s0.elapsed_time(s1) * 1e-3)
        for kk in (16, 256):                                       # K fused steps per launch
            agent._steps(kk)
            torch.cuda.synchronize(device)
            n_l = max(4096 // kk, 8)
            s0.record()
            for _ in range(n_l):
                agent._steps(kk)
            s1.record()
            torch.cuda.synchronize(device)
            extras["k%d_steps_per_sec" % kk] = n_l * kk / (s0.elapsed_time(s1) * 1e-3)
        eng.set_session(0 if args.no_session else 1)
        # the reference's whole inner loop (q_agent.py:174-189) with a greedy policy call per env transition:
        # train_frequency x (_policy -> add) + _step + loss read
        nloop = 2000
        agent._epsilon = 0.0
        states1 = s[:TRAIN_FREQUENCY * (nloop + 8)].reshape(-1, 1, D)
        def env_loop(n, off):
            for i in range(n):
                for j in range(TRAIN_FREQUENCY):
                    k = off + i * TRAIN_FREQUENCY + j
                    act = agent._policy(states1[k])
                    rb.add(s[k], act, r_py[k], s2[k], d_py[k])
                agent._step()
                eng.last_loss()
        env_loop(8, 0)
        c0 = time.perf_counter()
        env_loop(nloop, 8 * TRAIN_FREQUENCY)
        eng.synchronize()
        extras["env_loop_steps_per_sec"] = nloop / (time.perf_counter() - c0)
        extras["env_loop_note"] = "%d x (greedy Agent._policy + ReplayBuffer.add) + Agent._step + loss per step, wall clock" % TRAIN_FREQUENCY
        extras["replay_gather"] = bench_gather(torch, dqn_b200, eng, device, peak)
    except Exception as ex:       # extras never invalidate the main line
        extras["error"] = repr(ex)

    # ---- CPU baseline (oracle port on the host cores), bounded sample ---------------------------------
    rate, threads, done, dt = cpu_reference_steps_per_sec(20000, 50, budget_s=30.0)
    cpu = {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port",
           "sample": f"{done} oracle train steps (same workload) in {dt:.1f} s, BLAS threads={threads}, host cores={os.cpu_count()}"}

    line = {
        "metric": "train_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": secs_max / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": TRAIN_FREQUENCY * REC_BYTES_ALGO,
                "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "api": "ReplayBuffer.add x4 (host numpy) + Agent._step() + loss readback per step"
                       + ("" if args.no_session else "; Agent(session=True): commands served by the resident train-step kernel")},
        "gpu_launches": launches, "replay_samples_per_sec": value * B,
        "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_gather(torch, dqn_b200, eng, device, peak):
    """sample_batch as a standalone HBM-bound kernel: 65536 Philox-indexed samples per launch from the
    1M-slot ring into SoA outputs (reads 77 B + writes 77 B algorithmic per sample)."""
    import ctypes as C
    nb = 65536
    outs = [torch.empty(nb * D, dtype=torch.float32, device=device), torch.empty(nb, dtype=torch.int64, device=device),
            torch.empty(nb, dtype=torch.float32, device=device), torch.empty(nb * D, dtype=torch.float32, device=device),
            torch.empty(nb, dtype=torch.uint8, device=device)]
    lib, chk = eng.lib, dqn_b200.pkg._lib.check
    ptrs = [C.c_void_p(t.data_ptr()) for t in outs]
    for i in range(5):
        chk(lib.dqn_sample_batch_device(eng.h, 0, None, i, nb, *ptrs))
    flush_l2(torch, device)
    reps = 50
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(reps):
        chk(lib.dqn_sample_batch_device(eng.h, 0, None, 100 + i, nb, *ptrs))
    s1.record()
    torch.cuda.synchronize(device)
    dt = s0.elapsed_time(s1) * 1e-3 / reps
    gbs = nb * 2 * REC_BYTES_ALGO / dt / 1e9
    return {"samples_per_sec": nb / dt, "batch": nb, "us_per_launch": dt * 1e6, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / peak, "note": "ring (96 MB) fits in L2; see profiles/ for the >L2 ring measurement"}


def run_population(args):
    """BASELINE configs[2]: 1024 independent agents of the hyper-parameter sweep (per-agent gamma, batch in
    [38,70), Adam 1e-4, 40k-slot ring each), one CTA per agent, sharded over the ranks with NO data-path
    collective.  `value` = aggregate agent-train-steps/s; total work is fixed -> strong scaling."""
    import ctypes as C
    import torch
    import dqn_b200
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    n_global, ring = args.agents, 40_000                         # Test/lunar_lander_hyper_params.py:22
    pop = dqn_b200.Population(n_global, D, A, ring, dqn_b200.adam(1e-4), rank=rank, world_size=world, seed=1, device=local)
    eng, lib, chk = pop.engine, pop.engine.lib, dqn_b200.pkg._lib.check
    # synthetic transitions generated on the device (torch as RNG/allocator only), one block per agent
    g = torch.Generator(device=device)
    for i in range(pop.n_local):
        g.manual_seed(1000 + pop.global_id(i))
        s = torch.randn(ring, D, generator=g, device=device)
        s2 = torch.randn(ring, D, generator=g, device=device)
        r = 2.0 * torch.randn(ring, generator=g, device=device)
        a = torch.randint(0, A, (ring,), generator=g, device=device, dtype=torch.int64)
        d = (torch.rand(ring, generator=g, device=device) < 0.01).to(torch.uint8)
        chk(lib.dqn_store_device(eng.h, i, ring, C.c_void_p(s.data_ptr()), C.c_void_p(a.data_ptr()), C.c_void_p(r.data_ptr()),
                                 C.c_void_p(s2.data_ptr()), C.c_void_p(d.data_ptr())))
    torch.cuda.synchronize(device)
    kpl = args.steps_per_launch
    steps = max(kpl, (args.steps // kpl) * kpl)
    pop.train_steps(max(args.warmup, 3))
    flush_l2(torch, device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(steps // kpl):
            pop.train_steps(kpl)
        e1.record()
        torch.cuda.synchronize(device)
    secs = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs_max = float(t.item())
    if rank == 0:
        mean_b = float(np.mean([h["batch_size"] for h in dqn_b200.sweep_hparams(n_global)]))
        value = n_global * steps / secs_max
        sm_hz = (clk.summary()["sm_mhz"] or 1965) * 1e6
        fp32_peak = world * 148 * 128 * 2 * sm_hz / 1e9
        flops = value * mean_b * FLOP_PER_SAMPLE / 1e9
        peak, peak_src = measured_peaks()
        gbs = value * mean_b * REC_BYTES_ALGO / 1e9
        traffic = None
        try:      # DRAM bytes of one launch from the committed ncu capture, scaled to this rank's agents x steps per launch
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)["dqn_train_fused_kernel<4>/population"]
            traffic = tr["dram_bytes_per_launch"] / tr["agent_steps_per_launch"] * pop.n_local * kpl
        except Exception:
            pass
        print(json.dumps({
            "metric": "agent_train_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": secs_max / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: population of %d independent sweep agents, one CTA per agent, sharded over ranks, no collective" % n_global,
                       "obs_dim": D, "num_actions": A, "hidden": [32, 64], "ring_slots_per_agent": ring, "mean_batch": mean_b,
                       "optimizer": "adam(1e-4)", "steps_per_launch": kpl, "l2": "rings total %.1f GB per rank >> L2" % (pop.n_local * ring * 96 / 1e9)},
            "clocks": clk.summary(), "gpu_launches": steps // kpl,
            "replay_samples_per_sec": value * mean_b,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s", "frac": gbs / (peak * world), "traffic": traffic,
                         "kernel": "dqn_train_fused_kernel<4>", "peak_source": peak_src,
                         "note": "compute-bound on the fp32 pipe, not HBM: see fp32",
                         "fp32": {"achieved_gflops": flops, "ffma_peak_gflops": fp32_peak, "frac": flops / fp32_peak}}}))
    if world > 1:
        dist.destroy_process_group()


def run_replay(args):
    """Replay path alone, HBM-bound: 16M-slot ring (1.5 GB >> L2), 1M transitions per launch.  store = ReplayBuffer.add x 1M
    (SoA device arrays -> AoS ring), gather = sample_batch with Philox indices (ring -> the reference's five SoA arrays)."""
    import ctypes as C
    import torch
    import dqn_b200
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    ring, nb = 16 * 2**20, 2**20
    eng = dqn_b200.DqnEngine(D, A, ring, B, GAMMA, dqn_b200.adamw(LR), seed=0, device=0)
    lib, chk = eng.lib, dqn_b200.pkg._lib.check
    g = torch.Generator(device=device); g.manual_seed(0)
    src = [torch.randn(nb, D, generator=g, device=device), torch.randint(0, A, (nb,), generator=g, device=device, dtype=torch.int64),
           torch.randn(nb, generator=g, device=device), torch.randn(nb, D, generator=g, device=device),
           (torch.rand(nb, generator=g, device=device) < 0.01).to(torch.uint8)]
    out = [torch.empty_like(t) for t in src]
    sp, op = [C.c_void_p(t.data_ptr()) for t in src], [C.c_void_p(t.data_ptr()) for t in out]
    for _ in range(ring // nb):                                   # fill the ring once (also the warm-up of the store kernel)
        chk(lib.dqn_store_device(eng.h, 0, nb, *sp))
    for i in range(max(args.warmup, 3)):
        chk(lib.dqn_sample_batch_device(eng.h, 0, None, i, nb, *op))
    steps = max(min(args.steps, 200), 10)
    flush_l2(torch, device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with ClockSampler(0) as clk:
        e[0].record()
        for i in range(steps):
            chk(lib.dqn_sample_batch_device(eng.h, 0, None, 100 + i, nb, *op))
        e[1].record()
        for i in range(steps):
            chk(lib.dqn_store_device(eng.h, 0, nb, *sp))
        e[2].record()
        torch.cuda.synchronize(device)
    tg, ts = e[0].elapsed_time(e[1]) * 1e-3 / steps, e[1].elapsed_time(e[2]) * 1e-3 / steps
    peak, peak_src = measured_peaks()
    gbs_g, gbs_s = nb * 2 * REC_BYTES_ALGO / tg / 1e9, nb * 2 * REC_BYTES_ALGO / ts / 1e9
    print(json.dumps({
        "metric": "replay_samples_per_sec", "value": nb / tg, "unit": "samples/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": tg * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "replay path alone: 16M-slot ring (1.5 GB), 1M Philox-indexed samples per launch (sample_batch), 1M transitions per store",
                   "obs_dim": D, "l2": "ring 1.5 GB >> 126 MB L2; 512 MB flush before the timed region"},
        "clocks": clk.summary(), "gpu_launches": 2 * steps, "replay_stores_per_sec": nb / ts,
        "roofline": {"bound": "hbm", "achieved": gbs_g, "peak": peak, "unit": "GB/s", "frac": gbs_g / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "replay_gather_kernel", "algorithmic_bytes_per_launch": nb * 2 * REC_BYTES_ALGO,
                     "note": "77 B read + 77 B written per sample (reference dtypes); records are 96-byte AoS, so a random sample moves 3 sectors "
                             "= 96 B of DRAM for 77 B; store kernel: %.0f GB/s algorithmic (%.3f of peak)" % (gbs_s, gbs_s / peak)}}))


def run_episodes(args):
    """The reference's whole per-env-step loop (q_agent.py:174-203) for a population, on the device: epsilon-greedy policy ->
    (synthetic vectorised env) -> observe (store + episode bookkeeping + train gate) -> gated train step + hard sync.
    `value` = aggregate env steps/s; train steps happen on each agent's own train_frequency cadence (sweep draw, 2..15)."""
    import torch
    import dqn_b200
    from threadpoolctl import threadpool_limits
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    n_global, ring, T = args.agents, 40_000, 64
    pop = dqn_b200.Population(n_global, D, A, ring, dqn_b200.adam(1e-4), rank=rank, world_size=world, seed=1, device=local)
    n = pop.n_local
    pop.configure_episodes(max_episodes=10000, max_steps=1500, training_start=500, reward_to_reach=240.0)   # lunar_lander_hyper_params.py:22-30
    g = torch.Generator(device=device); g.manual_seed(77 + rank)
    obs = torch.randn(T, n, D, generator=g, device=device)
    rew = 2.0 * torch.randn(T, n, generator=g, device=device)
    done = (torch.rand(T, n, generator=g, device=device) < 0.01).to(torch.uint8)
    act = torch.empty(n, dtype=torch.int32, device=device)
    end = torch.empty(n, dtype=torch.uint8, device=device)

    def env_steps(k, t0):
        st = obs[(t0 - 1) % T]
        for t in range(t0, t0 + k):
            pop.policy(st, act)
            nxt = obs[t % T]
            pop.observe(st, act, rew[t % T], nxt, done[t % T], end)
            pop.train_flagged()
            st = nxt
    warm = 520                                                    # past training_start = 500: the train gate is live
    env_steps(warm, 0)
    torch.cuda.synchronize(device)
    steps = max(args.steps, 64)
    trained0 = sum(pop.engine.train_step_count(i) for i in range(n))
    flush_l2(torch, device)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        env_steps(steps, warm)
        e1.record()
        torch.cuda.synchronize(device)
    secs = e0.elapsed_time(e1) * 1e-3
    trained = sum(pop.engine.train_step_count(i) for i in range(n)) - trained0
    t = torch.tensor([secs, float(trained)], dtype=torch.float64, device=device)
    if world > 1:
        dist.barrier()
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        secs, trained = float(tm[0].item()), float(ts[1].item())
    if rank == 0:
        # CPU side: the restated reference loop (oracle) for one agent on one core, same cadence parameters
        from oracle import dqn_oracle as O
        from oracle.agent_oracle import OracleAgent
        from oracle.episode_oracle import EpisodeOracle
        hp = pop.hparams[0]
        rng = np.random.default_rng(0)
        theta = O.init_params(rng, D, A)
        oa = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-4), ring, D, hp["gamma"], hp["batch_size"], seed=1)
        eo = EpisodeOracle(oa, hp["epsilon"], hp["epsilon_decay_rate"], hp["min_epsilon"], 10000, 1500, 500, hp["train_frequency"],
                           hp["replace_frequency"], 240.0, A, seed=1)
        so, ro = rng.standard_normal((4096, D)).astype(np.float32), (2 * rng.standard_normal(4096)).astype(np.float32)
        with threadpool_limits(limits=1):
            for i in range(520):
                eo.observe(so[i], eo.policy(so[i])[0], ro[i], so[i + 1], False)
            c0 = time.perf_counter(); k = 0
            while time.perf_counter() - c0 < 10.0:
                i = 520 + k % 3000
                eo.observe(so[i], eo.policy(so[i])[0], ro[i], so[i + 1], False)
                k += 1
            cpu_rate = k / (time.perf_counter() - c0)
        print(json.dumps({
            "metric": "env_steps_per_sec", "value": n_global * steps / secs, "unit": "agent-env-steps/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "population of %d sweep agents, full device-side episode loop (policy -> observe -> gated train -> sync), "
                                   "synthetic vectorised env" % n_global, "obs_dim": D, "num_actions": A, "ring_slots_per_agent": ring,
                       "l2": "rings %.1f GB per rank >> L2" % (n * ring * 96 / 1e9)},
            "clocks": clk.summary(), "gpu_launches": int(steps * 3.5), "agent_train_steps_per_sec": trained / secs,
            "cpu_baseline": {"value": cpu_rate, "unit": "agent-env-steps/s", "cores": 1, "kind": "port",
                             "sample": "%d env steps of ONE agent (train_frequency %d) through oracle/episode_oracle.py in 10 s" % (k, hp["train_frequency"])}}))
    if world > 1:
        dist.destroy_process_group()


def run_dp(args):
    """BASELINE configs[3]: large-batch data-parallel DDQN, global batch 65536, hidden 1024x1024, D=8, A=4.
    Each rank: forward+backward on B/world rows -> ONE NCCL all-reduce of P+1 floats -> identical Adam.
    `value` = global train steps/s (strong scaling: the global batch is fixed)."""
    import ctypes as C
    import torch
    import dqn_b200
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    Bg, H, ring = args.batch, (args.hidden, args.hidden), 1_000_000
    tr = dqn_b200.LargeBatchTrainer(D, A, H, Bg, ring, GAMMA, dqn_b200.adamw(LR), rank=rank, world_size=world, seed=3,
                                    device=local, gemm_mode=args.gemm, collective=args.collective)
    params = dqn_b200.Model(A, hidden=H).init(np.random.default_rng(0), np.zeros((1, D), np.float32)) if False else None
    rng = np.random.default_rng(0)
    tree = {}
    for name, (fi, fo) in zip(dqn_b200.pkg.specs.MODULES, dqn_b200.pkg.specs.layer_shapes(D, A, H)):
        tree[name] = {"w": (rng.standard_normal((fi, fo)) / np.sqrt(fi)).clip(-2 / np.sqrt(fi), 2 / np.sqrt(fi)).astype(np.float32),
                      "b": np.zeros(fo, np.float32)}
    tr.set_params(tree, 0)
    tr.set_params(tree, 1)
    g = torch.Generator(device=device)
    g.manual_seed(1234)                                  # every rank holds the same ring replica
    s = torch.randn(ring, D, generator=g, device=device); s2 = torch.randn(ring, D, generator=g, device=device)
    r = 2.0 * torch.randn(ring, generator=g, device=device)
    a = torch.randint(0, A, (ring,), generator=g, device=device, dtype=torch.int64)
    d = (torch.rand(ring, generator=g, device=device) < 0.01).to(torch.uint8)
    tr.store_device(s, a, r, s2, d)
    del s, s2, r, a, d
    for _ in range(max(args.warmup, 3)):
        tr.step()
    flush_l2(torch, device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            tr.step()
        e1.record()
        torch.cuda.synchronize(device)
    secs = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs_max = float(t.item())
    loss = tr.loss()
    if rank == 0:
        flop_per_sample = 3 * 2 * (D * H[0] + H[0] * H[1] + H[1] * (1 + A)) + (2 * (D * H[0] + H[0] * H[1] + H[1] * (1 + A)) + 2 * (H[0] * H[1] + H[1] * (1 + A)))
        tflops = Bg * flop_per_sample * args.steps / secs_max / 1e12
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        mode_factor = {"fp32": None, "tc3xtf32": 6.0}[args.gemm]      # tf32 = 1/2 of bf16, 3 MMAs per product
        peak = pk["bf16_tflops_sustained"] * world
        print(json.dumps({
            "metric": "train_steps_per_sec", "value": args.steps / secs_max, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": secs_max / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if args.gemm == "fp32" else "f32 via 3xTF32 tensor-core split", "data": "synthetic",
            "config": {"workload": "configs[3]: large-batch data-parallel DDQN, global batch %d, hidden %dx%d" % (Bg, H[0], H[1]),
                       "obs_dim": D, "num_actions": A, "batch_local": Bg // world, "gemm": args.gemm,
                       "collective": {"p2p": "own kernel over NVLink peer memory (csrc/comm_p2p.cu), %d floats", "nccl": "NCCL all-reduce of %d floats",
                                      "none": "none (1 GPU), %d floats"}[tr.collective] % (tr.P + 1),
                       "l2": "activations %.1f GB per rank >> L2" % (3 * 2 * (Bg // world) * H[0] * 4 / 1e9)},
            # 18 kernels per step on every rank (index draw, gather, layer 1, 4 GEMMs, head, targets, 2 bias finishes, dh2, 3 partial
            # reductions, head grads, dW1, Adam) + the W2 transpose of the tensor-core mode + the peer-memory all-reduce when world > 1
            "clocks": clk.summary(),
            "gpu_launches": args.steps * (18 + (1 if args.gemm == "tc3xtf32" else 0) + (1 if tr.collective == "p2p" else 0)),
            "replay_samples_per_sec": Bg * args.steps / secs_max, "loss": loss,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak, "traffic": None,
                         "note": "peak = measured sustained dense bf16 (MEASURED_PEAKS.json); fp32-exact modes run at 1/%s of it at best (%s)"
                                 % ("n/a" if mode_factor is None else int(mode_factor), "FFMA pipe, 74 TF/GPU" if mode_factor is None else "tf32 = 1/2 bf16, x3 split"),
                         "flop_per_step": Bg * flop_per_sample}}))
    if world > 1:
        dist.destroy_process_group()


def run_per(args):
    """BASELINE configs[4]: prioritized-replay stress -- 16M-leaf sum tree, batch 32768: sample + priority update.
    No reference counterpart (the reference samples uniformly); one GPU (1.2 GB of ring + 134 MB tree fit)."""
    import torch
    import dqn_b200
    device = torch.device("cuda:0")
    cap, Bp = 16 * 2**20, 32768
    per = dqn_b200.PrioritizedSampler(cap, seed=1)
    g = torch.Generator(device=device); g.manual_seed(0)
    per.fill_device(torch.rand(cap, generator=g, device=device) + 1e-3)
    idx = torch.empty(Bp, dtype=torch.int64, device=device); pr = torch.empty(Bp, dtype=torch.float32, device=device)
    td = torch.randn(Bp, generator=g, device=device)
    def step(i):
        per.sample_device(i, idx, pr)
        per.update_device(idx, td, is_td=True)
    for i in range(max(args.warmup, 3)):
        step(i)
    flush_l2(torch, device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    steps = max(min(args.steps, 2000), 10)
    with ClockSampler(0) as clk:
        e[0].record()
        for i in range(steps):
            per.sample_device(1000 + i, idx, pr)
        e[1].record()
        for i in range(steps):
            per.update_device(idx, td, is_td=True)
        e[2].record()
        torch.cuda.synchronize(device)
    ts, tu = e[0].elapsed_time(e[1]) * 1e-3 / steps, e[1].elapsed_time(e[2]) * 1e-3 / steps
    peak, peak_src = measured_peaks()
    levels = 24
    gbs_s = Bp * levels * 4 / ts / 1e9
    gbs_u = Bp * (levels * 12 + 4) / tu / 1e9
    print(json.dumps({
        "metric": "per_samples_per_sec", "value": Bp / ts, "unit": "samples/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": (ts + tu) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: prioritized replay stress, 16M-leaf sum tree, batch 32768 (no reference counterpart)",
                   "l2": "tree 134 MB ~ L2 126 MB; 512 MB flush before the timed region"},
        "clocks": clk.summary(), "gpu_launches": steps * (1 + 1 + levels),
        "priority_updates_per_sec": Bp / tu, "us_per_sample_launch": ts * 1e6, "us_per_update": tu * 1e6,
        "roofline": {"bound": "hbm", "achieved": gbs_s, "peak": peak, "unit": "GB/s", "frac": gbs_s / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "per_sample_kernel", "note": "24 dependent 4-byte loads per sample: latency-bound pointer chase; update: %.1f GB/s algorithmic" % gbs_u}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200_000)
    ap.add_argument("--warmup", type=int, default=2_000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid (ncu): only the fused timed region, no e2e / extras / cpu baseline; not a bench value")
    ap.add_argument("--workload", default="single", choices=["single", "population", "dp", "per", "episodes", "replay"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tc3xtf32"])
    ap.add_argument("--agents", type=int, default=1024)
    ap.add_argument("--steps-per-launch", type=int, default=16)
    ap.add_argument("--no-session", action="store_true",
                    help="single workload: e2e through one launch per Agent._step() instead of the resident session kernel")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"],
                    help="dp workload: gradient all-reduce by the library's own peer-memory kernel (p2p) or by NCCL")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "cta", "cluster"],
                    help="train-step kernel of the single/population workloads: one CTA per agent, or one agent over a 4-CTA cluster "
                         "(auto = cluster while 4 * agents <= SMs)")
    args = ap.parse_args()
    os.environ["DQN_B200_STEP_KERNEL"] = args.step_kernel
    if args.impl == "reference":
        return run_reference(args)
    _, world, _ = dist_env()
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--workload", args.workload, "--agents", str(args.agents), "--steps-per-launch", str(args.steps_per_launch),
               "--batch", str(args.batch), "--hidden", str(args.hidden), "--gemm", args.gemm, "--step-kernel", args.step_kernel,
               "--collective", args.collective] + (["--no-session"] if args.no_session else [])
        sys.exit(subprocess.call(cmd))
    if args.workload == "population":
        return run_population(args)
    if args.workload == "dp":
        return run_dp(args)
    if args.workload == "per":
        return run_per(args)
    if args.workload == "episodes":
        return run_episodes(args)
    if args.workload == "replay":
        return run_replay(args)
    run_single(args)


if __name__ == "__main__":
    main()
