#!/usr/bin/env python
"""bench.py -- DQN train-steps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload default|single|population|dp|replay|per|episodes]

Main line = BASELINE.json configs[1]: single-agent fused dueling double-DQN train step, synthetic 1M-transition replay,
batch 64, D=8, A=4, hidden (32,64), AdamW(2e-4), gamma .99 on one B200.  A "step" is one train step (one minibatch:
sample -> targets -> loss -> backward -> Adam).  The single agent does not shard ("replicas only"), so
--gpus N runs N independent replicas, one process per GPU, no data-path collective; `value` = all ranks' steps /
max-over-ranks device time.

  value    steps/s with everything resident in HBM.  One "rep" = exactly --steps train steps (fused, at most 500 per
           persistent launch); reps are repeated -- L2 flushed before each -- until the timed region is >= 50 ms and the
           MEDIAN rep is reported (a 20-step rep is 110 us: one sample of it says nothing).
  e2e      steps/s through the reference-facing API (ReplayBuffer.add x train_frequency from host memory,
           Agent._step(), loss read back), 8000 steps (>= 50 ms) -- host<->device traffic inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md "Measurement"
  extras   (default workload) the sharded workloads AT THE SAME N, each a full line of its own (value, clocks, roofline):
           population = configs[2] (1024 sweep agents sharded over the ranks, no collective), dp = configs[3] (global batch
           65536, hidden 1024^2, tcgen05 3xTF32 GEMMs, gradient all-reduce), and on rank 0 replay (>L2 ring) and per
           (configs[4]); plus K=1/16/256 launch variants and the reference's whole inner loop.

`--impl reference` times the CPU path of the reference on the host cores: its own numba replay code (ReplayBuffer.add,
sample_batch -- imported from /root/reference when that tree exists, i.e. in the build container; the oracle's numba
restatement of the same two functions elsewhere) plus the NumPy oracle of targets / loss / backward / Adam (jax / haiku /
optax are not installable).  That and `cpu_baseline` are the only places bench.py executes oracle code.
"""
import argparse
import json
import math
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
STEPS_PER_LAUNCH = int(os.environ.get("DQN_BENCH_KPL", "500"))   # most train steps fused into one persistent launch
MIN_TIMED_S = 0.05                       # every timed region lasts at least this long (reps of the requested step count)
MAX_REPS = 2000
E2E_STEPS = 8000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1650.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled during the timed region (pynvml thread, 5 ms period)."""

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

    def _sample(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
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

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self.nv is not None and not self.samples:
            self._sample()
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


class Ctx:
    """Rank / device / process group of this process (one process per GPU under torchrun)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = self.dist = self.device = None

    def init_gpu(self):
        import torch
        self.torch = torch
        self.device = torch.device(f"cuda:{self.local}")
        torch.cuda.set_device(self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist
        return self

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def flush_l2(torch, device, _buf={}):
    """Write a 512 MB buffer (4x the 126 MB L2) on the current stream; enqueue only."""
    key = str(device)
    if key not in _buf:
        _buf[key] = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    _buf[key].fill_(1)


def timed_reps(ctx, rep_fn, min_seconds=MIN_TIMED_S, max_reps=MAX_REPS, flush=True, min_reps=1):
    """Run `rep_fn` (enqueues one rep of work on the current stream) until the timed region lasts >= min_seconds.
    Every rep is bracketed by its own CUDA events on the launching stream (the L2 flush sits between the brackets);
    barrier + synchronize on both sides of the whole region.  Returns (median rep seconds as the max over ranks, reps,
    total timed seconds on this rank, this rank's median, clocks)."""
    torch, device = ctx.torch, ctx.device
    ev = lambda: torch.cuda.Event(enable_timing=True)
    pilot = []                               # a few flushed reps, not reported: their median sizes the timed region
    for _ in range(5):
        a, b = ev(), ev()
        if flush:
            flush_l2(torch, device)
        a.record(); rep_fn(); b.record()
        torch.cuda.synchronize(device)
        pilot.append(a.elapsed_time(b) * 1e-3)
        if pilot[-1] > min_seconds:
            break
    est = ctx.max_over_ranks(float(np.median(pilot)))
    reps = int(min(max(math.ceil(1.1 * min_seconds / max(est, 1e-7)), min_reps, 1), max(max_reps, min_reps)))
    starts, ends = [ev() for _ in range(reps)], [ev() for _ in range(reps)]
    ctx.barrier()
    with ClockSampler(ctx.local) as clk:
        for i in range(reps):
            if flush:
                flush_l2(torch, device)
            starts[i].record()
            rep_fn()
            ends[i].record()
        torch.cuda.synchronize(device)
    ctx.barrier()
    times = np.array([s.elapsed_time(e) * 1e-3 for s, e in zip(starts, ends)])
    med_local = float(np.median(times))
    return ctx.max_over_ranks(med_local), reps, float(times.sum()), med_local, clk.summary()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path of Agent._step on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_sec(max_steps, warmup, budget_s, ring=N_RING, seed=0):
    """q_agent.py:146-169 on the CPU: train_frequency x ReplayBuffer.add + numba sample_batch (the reference's own module
    when /root/reference exists, else the oracle's numba restatement) + the NumPy oracle of preprocessing / targets /
    loss / grad / optax update.  numba JIT compile and warm-up excluded; BLAS threads 1 and all are both tried."""
    from threadpoolctl import threadpool_limits
    from oracle import dqn_oracle as O
    from oracle import replay_oracle as R
    rng = np.random.default_rng(seed)
    params = O.init_params(rng, D, A)
    target = O.tree_copy(params)
    opt, opt_state = O.OptSpec("adamw", LR), O.init_opt_state(params)
    ref = R.reference_replay_module()
    if ref is not None:
        rb, sample, replay_src = ref.ReplayBuffer(ring, (ring, D), (ring,)), ref.sample_batch, "reference module (General/Base/replay_buffer.py, numba)"
    else:
        rb, sample, replay_src = R.OracleReplay(ring, (ring, D), (ring,)), R.numba_sample_batch(), "oracle numba restatement of replay_buffer.py"
    s, a, r, s2, d = synthetic(rng, ring)
    a_py, r_py, d_py = a.tolist(), r.tolist(), d.tolist()
    for i in range(ring):                                        # ReplayBuffer.add x ring (replay_buffer.py:58-65)
        rb.add(s[i], a_py[i], r_py[i], s2[i], d_py[i])
    state = {"params": params, "opt": opt_state, "k": 0}

    def step():
        for _ in range(TRAIN_FREQUENCY):                         # q_agent.py:182
            k = state["k"] = (state["k"] + 1) % ring
            rb.add(s[k], a_py[k], r_py[k], s2[k], d_py[k])
        batch = sample(rb.size, rb.states, rb.actions, rb.rewards, rb.observations, rb.dones, B)   # q_agent.py:147-153
        state["params"], state["opt"] = O.train_step(state["params"], target, state["opt"], batch, GAMMA, opt)

    step()                                                       # numba compile
    best = None
    ncores = os.cpu_count() or 1
    for threads in sorted({1, ncores}):
        with threadpool_limits(limits=threads):
            for _ in range(max(warmup, 3)):
                step()
            t0 = time.perf_counter()
            done = 0
            while done < max_steps and time.perf_counter() - t0 < budget_s / 2:
                step()
                done += 1
            dt = time.perf_counter() - t0
        rate = done / dt
        if best is None or rate > best[0]:
            best = (rate, threads, done, dt)
    return best + (replay_src,)


def run_reference(args, ctx):
    if ctx.rank != 0:
        return None
    rate, threads, done, dt, src = cpu_reference_steps_per_sec(max(args.steps, 20000), args.warmup, budget_s=40.0)
    sample = (f"{done} CPU steps (4 x ReplayBuffer.add + numba sample_batch + NumPy train step; B=64, D=8, 1M-slot ring) "
              f"in {dt:.1f} s, BLAS threads={threads} (better of 1 and {os.cpu_count()}); replay half: {src}")
    return {
        "impl": "reference", "metric": "train_steps_per_sec", "value": rate, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": done, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "replay_samples_per_sec": rate * B,
        "note": "restated CPU reference (jax/haiku/optax not installable; the numba replay path is the reference's own "
                "code where /root/reference exists); host cores: %d" % (os.cpu_count() or 1),
    }


def workload_config():
    return {"workload": "configs[1]: single-agent fused dueling double-DQN train step, 1M-transition synthetic replay, batch 64",
            "obs_dim": D, "num_actions": A, "hidden": [32, 64], "batch": B, "ring_slots": N_RING, "gamma": GAMMA,
            "optimizer": "adamw(2e-4, wd 1e-4)", "steps_per_launch": STEPS_PER_LAUNCH,
            "step_kernel": os.environ.get("DQN_B200_STEP_KERNEL", "auto"),
            "multi_gpu": "replicas only (single agent does not shard); the sharded workloads at this N are in extras",
            "l2": "ring 96 MB < 126 MB L2: L2 flushed (512 MB write) before every timed rep; every step gathers 64 random, mostly first-touch records"}


# ---------------------------------------------------------------------------------------------------
# configs[1]: single agent
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


def run_single(args, ctx, with_extras=True):
    import dqn_b200
    torch, device = ctx.torch, ctx.device
    agent, data = build_agent(dqn_b200, ctx.local, seed=ctx.rank, session=not args.no_session)
    eng = agent._engine
    steps = args.steps

    # ---- value: device-resident, fused ----------------------------------------------------------
    def rep():
        left = steps
        while left > 0:
            k = min(left, STEPS_PER_LAUNCH)
            agent._steps(k)
            left -= k
    launches_per_rep = -(-steps // STEPS_PER_LAUNCH)
    agent._steps(max(args.warmup, 3))
    rep_s, reps, timed_s, rep_local, clocks = timed_reps(ctx, rep)
    value = ctx.world * steps / rep_s

    if args.profile:
        ctx.barrier()
        return {"profile_only": True, "steps_per_sec": value, "launches": reps * launches_per_rep} if ctx.rank == 0 else None

    # ---- e2e: reference-facing API, host buffers, copies inside the timed region ------------------
    e2e_steps = E2E_STEPS
    nloop = 2000
    rb = agent._replay_buffer
    nhost = TRAIN_FREQUENCY * (max(e2e_steps, nloop) + 16)
    s, a, r, s2, d = [x[:nhost] for x in data]
    a_py, r_py, d_py = a.tolist(), r.tolist(), d.tolist()

    def e2e_loop(n, off=0):
        # Software-pipelined the way the reference's loop allows: _step() only ENQUEUES step i (a command to the resident
        # kernel, which holds two command slots), so the host stages the train_frequency add()s of step i+1 while the device
        # works, publishes step i+1 and THEN reads step i's loss (dqn_get_loss_lagged).  Every step's loss is read back, one
        # read per step, one step behind, inside the timed region; the device never waits for the host between two steps.
        last = 0.0
        for i in range(n):
            for j in range(TRAIN_FREQUENCY):                      # q_agent.py:182  one add() per env transition
                k = off + i * TRAIN_FREQUENCY + j
                rb.add(s[k], a_py[k], r_py[k], s2[k], d_py[k])
            agent._step()                                          # q_agent.py:187
            if i:
                last = rb.last_loss(1)                             # device -> host read of the previous step's loss
        return rb.last_loss()

    e2e_loop(8)
    eng.synchronize()                       # (session mode: retire the resident kernel before other work uses the stream)
    flush_l2(torch, device)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps, off=8 * TRAIN_FREQUENCY)
    eng.synchronize()                       # inside the timed region: the session's state is back in HBM
    e1.record()
    torch.cuda.synchronize(device)
    e2e_wall = time.perf_counter() - t0
    e2e_secs = ctx.max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, e2e_wall))
    e2e_value = ctx.world * e2e_steps / e2e_secs
    ctx.barrier()

    extras = {}
    if ctx.rank == 0 and with_extras:
        single_extras(args, ctx, dqn_b200, agent, (s, a_py, r_py, s2, d_py), nloop, extras)
    eng.synchronize()
    agent._engine.close()
    del agent, eng, rb
    if with_extras:                          # every rank takes part in the sharded workloads
        sharded_extras(args, ctx, extras)
    if ctx.rank != 0:
        return None

    # ---- roofline of the dominant kernel, measured live ---------------------------------------------
    pk, peak_src = measured_peaks()
    peak = float(pk["hbm_gbs"])
    launch_s = rep_local / launches_per_rep
    steps_per_launch = steps / launches_per_rep
    algo_bytes = steps_per_launch * (B * REC_BYTES_ALGO + 4)          # gathered records + one loss store per step
    achieved = algo_bytes / launch_s / 1e9
    sm_hz = (clocks["sm_mhz"] or 1965) * 1e6
    fp32_one_sm = 128 * 2 * sm_hz / 1e9
    flops = B * FLOP_PER_SAMPLE / (rep_local / steps) / 1e9
    cluster = args.step_kernel in ("auto", "cluster")
    n_sm = 4 if cluster else 1
    kname = "dqn_train_cluster_kernel<4>" if cluster else "dqn_train_fused_kernel<4>"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_from_profile(kname, "steps_per_launch", steps_per_launch), "kernel": kname, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": launch_s * 1e3,
                "note": "latency-bound by construction: one agent's steps are a serial chain (step t+1 needs theta_t) on %s; "
                        "theta/theta^-/grads stay in shared memory, so the only HBM traffic is 64 gathered records per step "
                        "(prefetched one step ahead)" % ("a 4-CTA cluster (4 of 148 SMs)" if cluster else "one CTA (1 of 148 SMs)"),
                "fp32": {"achieved_gflops": flops, "sms_used": n_sm, "ffma_peak_gflops_of_sms_used": n_sm * fp32_one_sm,
                         "frac_of_sms_used": flops / (n_sm * fp32_one_sm), "flop_per_step": B * FLOP_PER_SAMPLE}}

    # ---- CPU baseline on the host cores, bounded sample (rank 0, N = 1 only) ---------------------------
    cpu = None
    if ctx.world == 1 and not args.no_cpu_baseline:
        rate, threads, done, dt, src = cpu_reference_steps_per_sec(20000, 50, budget_s=24.0)
        cpu = {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"{done} CPU steps of the same workload (4 x ReplayBuffer.add + numba sample_batch + NumPy train step) in {dt:.1f} s, "
                         f"BLAS threads={threads} (better of 1 and {os.cpu_count()}), host cores={os.cpu_count()}; replay half: {src}"}

    return {
        "metric": "train_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": rep_s / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "clocks": clocks,
        "timing": {"reps": reps, "steps_per_rep": steps, "launches_per_rep": launches_per_rep, "statistic": "median rep, max over ranks",
                   "timed_region_s": timed_s, "min_timed_region_s": MIN_TIMED_S},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": TRAIN_FREQUENCY * REC_BYTES_ALGO,
                "d2h_bytes_per_step": 4, "steps": e2e_steps, "timed_region_s": e2e_secs,
                "loss_read_lag_steps": 1,
                "api": "ReplayBuffer.add x4 (host numpy) + Agent._step() + loss readback per step, one step behind (step i+1 is staged and "
                       "published while step i runs, then step i's loss is read: dqn_get_loss_lagged)"
                       + ("" if args.no_session else "; Agent(session=True): commands served by the resident train-step kernel")},
        "gpu_launches": reps * launches_per_rep, "replay_samples_per_sec": value * B,
        "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
    }


def traffic_from_profile(kname, per_key, units):
    """DRAM bytes of one launch from the committed ncu --set full capture, scaled to this run's work per launch."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f)[kname]
            return tr["dram_bytes_per_launch"] / tr[per_key] * units
        except Exception:
            continue
    return None


def guarded(extras, key, fn):
    """One failing extra never drops the others (nor the main line)."""
    try:
        extras[key] = fn()
    except Exception as ex:
        extras[key] = {"error": repr(ex)}


def single_extras(args, ctx, dqn_b200, agent, host, nloop, extras):
    """Rank 0: launch-count variants of the single-agent step and the reference's whole inner loop."""
    torch, device = ctx.torch, ctx.device
    eng, rb = agent._engine, agent._replay_buffer
    s, a_py, r_py, s2, d_py = host
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def k_variant(kk):
        def run():
            eng.set_session(0)                                     # launches, not the resident kernel
            agent._steps(kk)
            torch.cuda.synchronize(device)
            n_l = max(8192 // kk, 16)
            s0.record()
            for _ in range(n_l):
                if kk == 1:
                    agent._step()                                  # one kernel launch per Agent._step()
                else:
                    agent._steps(kk)
            s1.record()
            torch.cuda.synchronize(device)
            return n_l * kk / (s0.elapsed_time(s1) * 1e-3)
        return run
    for kk in (1, 16, 256):                                        # K fused steps per launch
        guarded(extras, "k%d_steps_per_sec" % kk, k_variant(kk))

    def env_loop_run():
        # the reference's whole inner loop (q_agent.py:174-189) with a greedy policy call per env transition:
        # train_frequency x (_policy -> add) + _step + loss read
        eng.set_session(0 if args.no_session else 1)
        agent._epsilon = 0.0
        states1 = s.reshape(-1, 1, D)

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
        return {"value": nloop / (time.perf_counter() - c0), "unit": "steps/s", "steps": nloop,
                "note": "%d x (greedy Agent._policy + ReplayBuffer.add) + Agent._step + loss per step, wall clock" % TRAIN_FREQUENCY}
    guarded(extras, "env_loop", env_loop_run)
    pk, _ = measured_peaks()
    guarded(extras, "replay_gather_l2_resident", lambda: bench_gather(torch, dqn_b200, eng, device, float(pk["hbm_gbs"])))


def sharded_extras(args, ctx, extras):
    """Every rank: the workloads that DO shard, at this N, each under its own guard; rank 0 alone: replay, per."""
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup, sub.agents, sub.steps_per_launch = 256, 3, 1024, 128
    for key, fn in (("population", run_population),):
        try:
            line = fn(sub, ctx)
        except Exception as ex:
            line = {"error": repr(ex)}
        if ctx.rank == 0:
            extras[key] = line
        ctx.torch.cuda.empty_cache()
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup, sub.gemm, sub.collective, sub.batch, sub.hidden = 20, 3, "tc3xtf32", "auto", 65536, 1024
    try:
        line = run_dp(sub, ctx)
    except Exception as ex:
        line = {"error": repr(ex)}
    if ctx.rank == 0:
        extras["dp"] = line
    ctx.torch.cuda.empty_cache()
    if ctx.rank == 0:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = 600, 3
        guarded(extras, "replay", lambda: run_replay(sub, ctx))
        ctx.torch.cuda.empty_cache()
        sub.steps = 2000
        guarded(extras, "per", lambda: run_per(sub, ctx))
        ctx.torch.cuda.empty_cache()
    ctx.barrier()


def bench_gather(torch, dqn_b200, eng, device, peak):
    """sample_batch as a standalone kernel: 65536 Philox-indexed samples per launch from the (L2-resident) 1M-slot ring
    into SoA outputs (reads 77 B + writes 77 B algorithmic per sample); the HBM-bound figure is extras.replay."""
    import ctypes as C
    nb = 65536
    outs = [torch.empty(nb * D, dtype=torch.float32, device=device), torch.empty(nb, dtype=torch.int64, device=device),
            torch.empty(nb, dtype=torch.float32, device=device), torch.empty(nb * D, dtype=torch.float32, device=device),
            torch.empty(nb, dtype=torch.uint8, device=device)]
    lib, chk = eng.lib, dqn_b200.pkg._lib.check
    ptrs = [C.c_void_p(t.data_ptr()) for t in outs]
    eng.set_session(0)
    for i in range(5):
        chk(lib.dqn_sample_batch_device(eng.h, 0, None, i, nb, *ptrs))
    reps = 200
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(reps):
        chk(lib.dqn_sample_batch_device(eng.h, 0, None, 100 + i, nb, *ptrs))
    s1.record()
    torch.cuda.synchronize(device)
    dt = s0.elapsed_time(s1) * 1e-3 / reps
    gbs = nb * 2 * REC_BYTES_ALGO / dt / 1e9
    return {"samples_per_sec": nb / dt, "batch": nb, "us_per_launch": dt * 1e6, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / peak, "note": "ring (96 MB) fits in L2"}


# ---------------------------------------------------------------------------------------------------
# configs[2]: population of sweep agents, sharded, no collective
# ---------------------------------------------------------------------------------------------------
def run_population(args, ctx):
    """1024 independent agents of the hyper-parameter sweep (per-agent gamma, batch in [38,70], Adam 1e-4, 40k-slot ring
    each), sharded over the ranks with NO data-path collective.  `value` = aggregate agent-train-steps/s; total work is
    fixed -> strong scaling.  One rep = `steps` train steps of every agent (steps_per_launch per launch)."""
    import ctypes as C
    import dqn_b200
    torch, device = ctx.torch, ctx.device
    n_global, ring = args.agents, 40_000                         # Test/lunar_lander_hyper_params.py:22
    pop = dqn_b200.Population(n_global, D, A, ring, dqn_b200.adam(1e-4), rank=ctx.rank, world_size=ctx.world, seed=1, device=ctx.local)
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
    kpl = max(1, min(args.steps_per_launch, args.steps))
    steps = max(kpl, (args.steps // kpl) * kpl)
    pop.train_steps(max(args.warmup, 3))

    def rep():
        for _ in range(steps // kpl):
            pop.train_steps(kpl)
    rep_s, reps, timed_s, rep_local, clocks = timed_reps(ctx, rep)
    hp_all = dqn_b200.sweep_hparams(n_global)
    local_tiles = sum((hp["batch_size"] + 63) // 64 for hp in pop.hparams)
    max_tiles = ctx.max_over_ranks(local_tiles)
    # every rank's parameters after the run, hashed: a sharded run must reproduce the unsharded one agent by agent
    digest = float(np.frombuffer(pop.params_flat(0).tobytes()[:8], dtype=np.uint32)[0] & 0xFFFFFF)
    digest_sum = ctx.sum_over_ranks(digest)
    n_local = pop.n_local
    pop.engine.close()
    del pop, eng
    if ctx.rank != 0:
        return None
    pop_kernel = "dqn_train_fused_kernel<4>" if args.step_kernel == "cta" else "dqn_train_tc_kernel<4>"
    mean_b = float(np.mean([h["batch_size"] for h in hp_all]))
    two_tile = float(np.mean([h["batch_size"] > 64 for h in hp_all]))
    value = n_global * steps / rep_s
    sm_hz = (clocks["sm_mhz"] or 1965) * 1e6
    fp32_peak = ctx.world * 148 * 128 * 2 * sm_hz / 1e9
    flops = value * mean_b * FLOP_PER_SAMPLE / 1e9
    pk, peak_src = measured_peaks()
    peak = float(pk["hbm_gbs"])
    gbs = value * mean_b * REC_BYTES_ALGO / 1e9
    return {
        "metric": "agent_train_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": rep_s / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: population of %d independent sweep agents, one CTA per agent, sharded over ranks, no collective" % n_global,
                   "obs_dim": D, "num_actions": A, "hidden": [32, 64], "ring_slots_per_agent": ring, "mean_batch": mean_b,
                   "agents_with_batch_over_64": two_tile, "agents_per_rank": n_local, "tiles_on_busiest_rank": max_tiles,
                   "optimizer": "adam(1e-4)", "steps_per_launch": kpl,
                   "step_kernel": "cta (fp32 FFMA)" if args.step_kernel == "cta" else "cta_tc (layer-2 products on tcgen05 3xTF32, 512 threads)", "l2": "rings total %.1f GB per rank >> L2; flushed before every rep" % (n_local * ring * 96 / 1e9)},
        "clocks": clocks, "gpu_launches": reps * (steps // kpl),
        "timing": {"reps": reps, "steps_per_rep": steps, "statistic": "median rep, max over ranks", "timed_region_s": timed_s},
        "replay_samples_per_sec": value * mean_b, "param_digest_sum": digest_sum,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * ctx.world, "unit": "GB/s", "frac": gbs / (peak * ctx.world),
                     "traffic": traffic_from_profile(pop_kernel + "/population", "agent_steps_per_launch", n_local * kpl),
                     "kernel": pop_kernel, "peak_source": peak_src,
                     "note": "latency / issue-bound on the SM (fp32 pipe + tcgen05 round trips), not HBM: see fp32 (algorithmic flops, whichever pipe runs them)",
                     "fp32": {"achieved_gflops": flops, "ffma_peak_gflops": fp32_peak, "frac": flops / fp32_peak}}}


# ---------------------------------------------------------------------------------------------------
# replay path alone (HBM-bound)
# ---------------------------------------------------------------------------------------------------
def run_replay(args, ctx):
    """16M-slot ring (1.5 GB >> L2), 1M transitions per launch.  store = ReplayBuffer.add x 1M (SoA device arrays -> AoS
    ring), gather = sample_batch with Philox indices (ring -> the reference's five SoA arrays)."""
    import ctypes as C
    import dqn_b200
    torch, device = ctx.torch, ctx.device
    ring, nb = 16 * 2**20, 2**20
    eng = dqn_b200.DqnEngine(D, A, ring, B, GAMMA, dqn_b200.adamw(LR), seed=0, device=ctx.local)
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
    steps = max(min(args.steps, 2000), 10)
    flush_l2(torch, device)
    torch.cuda.synchronize(device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with ClockSampler(ctx.local) as clk:
        e[0].record()
        for i in range(steps):
            chk(lib.dqn_sample_batch_device(eng.h, 0, None, 100 + i, nb, *op))
        e[1].record()
        for i in range(steps):
            chk(lib.dqn_store_device(eng.h, 0, nb, *sp))
        e[2].record()
        torch.cuda.synchronize(device)
    tg, ts = e[0].elapsed_time(e[1]) * 1e-3 / steps, e[1].elapsed_time(e[2]) * 1e-3 / steps
    eng.close()
    pk, peak_src = measured_peaks()
    peak = float(pk["hbm_gbs"])
    gbs_g, gbs_s = nb * 2 * REC_BYTES_ALGO / tg / 1e9, nb * 2 * REC_BYTES_ALGO / ts / 1e9
    return {
        "metric": "replay_samples_per_sec", "value": nb / tg, "unit": "samples/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": tg * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "replay path alone: 16M-slot ring (1.5 GB), 1M Philox-indexed samples per launch (sample_batch), 1M transitions per store",
                   "obs_dim": D, "l2": "ring 1.5 GB >> 126 MB L2 (every launch streams 100+ MB of random records); 512 MB flush before the timed region"},
        "clocks": clk.summary(), "gpu_launches": 2 * steps, "replay_stores_per_sec": nb / ts,
        "timing": {"timed_region_s": (tg + ts) * steps},
        "roofline": {"bound": "hbm", "achieved": gbs_g, "peak": peak, "unit": "GB/s", "frac": gbs_g / peak, "peak_source": peak_src,
                     "traffic": traffic_from_profile("replay_gather_kernel", "samples_per_launch", nb),
                     "kernel": "replay_gather_kernel", "algorithmic_bytes_per_launch": nb * 2 * REC_BYTES_ALGO,
                     "store": {"achieved": gbs_s, "frac": gbs_s / peak, "kernel": "replay_store_kernel"},
                     "note": "77 B read + 77 B written per sample (reference dtypes); records are 96-byte AoS, so a random sample moves two 64-byte "
                             "DRAM atoms = 128 B for 77 B"}}


# ---------------------------------------------------------------------------------------------------
# episode loop of a population on the device
# ---------------------------------------------------------------------------------------------------
def run_episodes(args, ctx):
    """The reference's whole per-env-step loop (q_agent.py:174-203) for a population, on the device: epsilon-greedy policy ->
    (synthetic vectorised env) -> observe (store + episode bookkeeping + train gate) -> gated train step + hard sync.
    `value` = aggregate env steps/s; train steps happen on each agent's own train_frequency cadence (sweep draw, 2..15)."""
    import dqn_b200
    from threadpoolctl import threadpool_limits
    torch, device = ctx.torch, ctx.device
    n_global, ring, T = args.agents, 40_000, 64
    pop = dqn_b200.Population(n_global, D, A, ring, dqn_b200.adam(1e-4), rank=ctx.rank, world_size=ctx.world, seed=1, device=ctx.local)
    n = pop.n_local
    pop.configure_episodes(max_episodes=10000, max_steps=1500, training_start=500, reward_to_reach=240.0)   # lunar_lander_hyper_params.py:22-30
    g = torch.Generator(device=device); g.manual_seed(77 + ctx.rank)
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
    steps = max(args.steps, 1024)
    trained0 = sum(pop.engine.train_step_count(i) for i in range(n))
    flush_l2(torch, device)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx.local) as clk:
        e0.record()
        env_steps(steps, warm)
        e1.record()
        torch.cuda.synchronize(device)
    secs = ctx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    trained = ctx.sum_over_ranks(sum(pop.engine.train_step_count(i) for i in range(n)) - trained0)
    ctx.barrier()
    if ctx.rank != 0:
        return None
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
    return {
        "metric": "env_steps_per_sec", "value": n_global * steps / secs, "unit": "agent-env-steps/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": warm, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "population of %d sweep agents, full device-side episode loop (policy -> observe -> gated train -> sync), "
                               "synthetic vectorised env" % n_global, "obs_dim": D, "num_actions": A, "ring_slots_per_agent": ring,
                   "l2": "rings %.1f GB per rank >> L2" % (n * ring * 96 / 1e9)},
        "clocks": clk.summary(), "gpu_launches": int(steps * 3.5), "agent_train_steps_per_sec": trained / secs,
        "cpu_baseline": {"value": cpu_rate, "unit": "agent-env-steps/s", "cores": 1, "kind": "port",
                         "sample": "%d env steps of ONE agent (train_frequency %d) through oracle/episode_oracle.py in 10 s" % (k, hp["train_frequency"])}}


# ---------------------------------------------------------------------------------------------------
# configs[3]: large-batch data-parallel step
# ---------------------------------------------------------------------------------------------------
def run_dp(args, ctx):
    """Large-batch data-parallel DDQN, global batch 65536, hidden 1024x1024, D=8, A=4.  Each rank: forward+backward on
    B/world rows -> ONE all-reduce of P+1 floats -> identical Adam.  `value` = global train steps/s (strong scaling:
    the global batch is fixed).  One rep = one step; the median of max(steps, 50 ms worth) reps is reported."""
    import hashlib
    import dqn_b200
    torch, device = ctx.torch, ctx.device
    Bg, H, ring = args.batch, (args.hidden, args.hidden), 1_000_000
    tr = dqn_b200.LargeBatchTrainer(D, A, H, Bg, ring, GAMMA, dqn_b200.adamw(LR), rank=ctx.rank, world_size=ctx.world, seed=3,
                                    device=ctx.local, gemm_mode=args.gemm, collective=args.collective)
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
    rep_s, reps, timed_s, rep_local, clocks = timed_reps(ctx, tr.step, min_reps=args.steps, max_reps=max(args.steps, 200))
    loss = tr.loss()
    # replicas must stay bit-identical: hash every rank's parameters and compare
    flat = np.concatenate([np.concatenate([v["w"].ravel(), v["b"].ravel()]) for v in tr.get_params(0).values()])
    digest = int.from_bytes(hashlib.sha1(flat.tobytes()).digest()[:6], "little")
    digests = [digest]
    if ctx.dist is not None:
        t = torch.tensor([digest], dtype=torch.int64, device=device)
        allt = [torch.zeros_like(t) for _ in range(ctx.world)]
        ctx.dist.all_gather(allt, t)
        digests = [int(x.item()) for x in allt]
    collective, P = tr.collective, tr.P
    tr.close()
    del tr
    if ctx.rank != 0:
        return None
    flop_per_sample = 3 * 2 * (D * H[0] + H[0] * H[1] + H[1] * (1 + A)) + (2 * (D * H[0] + H[0] * H[1] + H[1] * (1 + A)) + 2 * (H[0] * H[1] + H[1] * (1 + A)))
    tflops = Bg * flop_per_sample / rep_s / 1e12
    pk, peak_src = measured_peaks()
    mode_factor = {"fp32": None, "tc3xtf32": 6.0}[args.gemm]      # tf32 = 1/2 of bf16, 3 MMAs per product
    peak = float(pk["bf16_tflops_sustained"]) * ctx.world
    launches_per_step = 17 + (1 if args.gemm == "tc3xtf32" else 0) + (2 if collective == "p2p" else 0)
    return {
        "metric": "train_steps_per_sec", "value": 1.0 / rep_s, "unit": "steps/s", "n_gpus": ctx.world, "steps": reps,
        "warmup": max(args.warmup, 3), "ms_per_step": rep_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if args.gemm == "fp32" else "f32 via 3xTF32 tensor-core split", "data": "synthetic",
        "config": {"workload": "configs[3]: large-batch data-parallel DDQN, global batch %d, hidden %dx%d" % (Bg, H[0], H[1]),
                   "obs_dim": D, "num_actions": A, "batch_local": Bg // ctx.world, "gemm": args.gemm,
                   "collective": {"p2p": "own kernel over NVLink peer memory (csrc/comm_p2p.cu), %d floats", "nccl": "NCCL all-reduce of %d floats",
                                  "none": "none (1 GPU), %d floats"}[collective] % (P + 1),
                   "l2": "activations %.1f GB per rank >> L2; flushed before every step" % (3 * 2 * (Bg // ctx.world) * H[0] * 4 / 1e9)},
        "clocks": clocks, "gpu_launches": reps * launches_per_step,
        "timing": {"reps": reps, "steps_per_rep": 1, "statistic": "median step, max over ranks", "timed_region_s": timed_s},
        "replay_samples_per_sec": Bg / rep_s, "loss": loss,
        "replicas": {"param_digests": digests, "identical": len(set(digests)) == 1},
        "roofline": {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak, "traffic": None,
                     "peak_source": peak_src,
                     "frac_of_mode_ceiling": None if mode_factor is None else tflops / (peak / mode_factor),
                     "note": "peak = measured sustained dense bf16 (MEASURED_PEAKS.json); fp32-exact modes run at 1/%s of it at best (%s)"
                             % ("n/a" if mode_factor is None else int(mode_factor), "FFMA pipe, 74 TF/GPU" if mode_factor is None else "tf32 = 1/2 bf16, x3 split"),
                     "flop_per_step": Bg * flop_per_sample}}


# ---------------------------------------------------------------------------------------------------
# configs[4]: prioritized replay stress (no reference counterpart)
# ---------------------------------------------------------------------------------------------------
def run_per(args, ctx):
    """16M-leaf sum tree, batch 32768: sample + priority update.  No reference counterpart (the reference samples
    uniformly); one GPU (1.2 GB of ring + 134 MB tree fit)."""
    import dqn_b200
    torch, device = ctx.torch, ctx.device
    cap, Bp = 16 * 2**20, 32768
    per = dqn_b200.PrioritizedSampler(cap, seed=1, device=ctx.local)
    g = torch.Generator(device=device); g.manual_seed(0)
    per.fill_device(torch.rand(cap, generator=g, device=device) + 1e-3)
    idx = torch.empty(Bp, dtype=torch.int64, device=device); pr = torch.empty(Bp, dtype=torch.float32, device=device)
    td = torch.randn(Bp, generator=g, device=device)
    for i in range(max(args.warmup, 3)):
        per.sample_device(i, idx, pr)
        per.update_device(idx, td, is_td=True)
    flush_l2(torch, device)
    torch.cuda.synchronize(device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    steps = max(min(args.steps, 2000), 10)
    with ClockSampler(ctx.local) as clk:
        e[0].record()
        for i in range(steps):
            per.sample_device(1000 + i, idx, pr)
        e[1].record()
        for i in range(steps):
            per.update_device(idx, td, is_td=True)
        e[2].record()
        torch.cuda.synchronize(device)
    ts, tu = e[0].elapsed_time(e[1]) * 1e-3 / steps, e[1].elapsed_time(e[2]) * 1e-3 / steps
    pk, peak_src = measured_peaks()
    peak = float(pk["hbm_gbs"])
    levels = 24
    gbs_s = Bp * levels * 4 / ts / 1e9
    gbs_u = Bp * (levels * 12 + 4) / tu / 1e9
    launches_per_update = getattr(per, "launches_per_update", 1 + levels)
    return {
        "metric": "per_samples_per_sec", "value": Bp / ts, "unit": "samples/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": (ts + tu) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: prioritized replay stress, 16M-leaf sum tree, batch 32768 (no reference counterpart)",
                   "l2": "tree 134 MB ~ L2 126 MB; 512 MB flush before the timed region"},
        "clocks": clk.summary(), "gpu_launches": steps * (1 + launches_per_update),
        "timing": {"timed_region_s": (ts + tu) * steps},
        "priority_updates_per_sec": Bp / tu, "us_per_sample_launch": ts * 1e6, "us_per_update": tu * 1e6,
        "roofline": {"bound": "hbm", "achieved": gbs_s, "peak": peak, "unit": "GB/s", "frac": gbs_s / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "per_sample_kernel", "note": "24 dependent 4-byte loads per sample: latency-bound pointer chase; update: %.1f GB/s algorithmic" % gbs_u}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid (ncu): only the fused timed region, no e2e / extras / cpu baseline; not a bench value")
    ap.add_argument("--workload", default="default", choices=["default", "single", "population", "dp", "per", "episodes", "replay"],
                    help="default = the single-agent line with the sharded workloads at the same N embedded in extras")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--gemm", default="tc3xtf32", choices=["fp32", "tc3xtf32"])
    ap.add_argument("--agents", type=int, default=1024)
    ap.add_argument("--steps-per-launch", type=int, default=128)
    ap.add_argument("--no-session", action="store_true",
                    help="single workload: e2e through one launch per Agent._step() instead of the resident session kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"],
                    help="dp workload: gradient all-reduce by the library's own peer-memory kernel (p2p) or by NCCL")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "cta", "cluster", "cta_tc"],
                    help="train-step kernel of the single/population workloads: one CTA per agent, or one agent over a 4-CTA cluster "
                         "(auto = cluster while 4 * agents <= SMs)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner, ...) are sent to stderr
    # for the whole run; the line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    emit = lambda line: (real_stdout.write(json.dumps(line) + "\n"), real_stdout.flush())
    defaults = {"default": (200_000, 2_000), "single": (200_000, 2_000), "population": (256, 3), "dp": (20, 3), "per": (2000, 3),
                "episodes": (1024, 3), "replay": (600, 3)}[args.workload]
    if args.steps is None:
        args.steps = defaults[0]
    if args.warmup is None:
        args.warmup = defaults[1]
    os.environ["DQN_B200_STEP_KERNEL"] = args.step_kernel
    ctx = Ctx()
    if args.impl == "reference":
        line = run_reference(args, ctx)
        if line is not None:
            emit(line)
        return
    if args.gpus != ctx.world and ctx.world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    ctx.init_gpu()
    try:
        fn = {"default": lambda a, c: run_single(a, c, True), "single": lambda a, c: run_single(a, c, False),
              "population": run_population, "dp": run_dp, "per": run_per, "episodes": run_episodes, "replay": run_replay}[args.workload]
        line = fn(args, ctx)
        if ctx.rank == 0 and line is not None:
            emit(line)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
