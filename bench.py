#!/usr/bin/env python3
"""bench.py — headline benchmark of the DQN-MARL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Metric (BASELINE.json): env agent-steps/s.  Workload at N=1 = BASELINE.json configs[1] ("CA-dqn1
single-room cellular-automaton evacuation, 4096 batched envs, env-step-only throughput on 1 B200"):
36x30 room, 150 people per env, 4096 envs, uniform random robot actions, auto-reset.  One "step" = one
fused env-step launch over the whole batch = 4096*150 agent-steps (every person counted every step,
SURVEY.md §8d).

  value     device-resident throughput: K back-to-back steps, inputs already in HBM, CUDA events.
            The whole state of one batch (27 MB) would sit in the 126 MB L2, so the timed loop ROTATES over
            enough independent batches that the state touched between two visits of the same batch exceeds
            L2 (config.l2 says how many) — every step reads its state from HBM.
  e2e       same metric through the host-facing call: actions from pinned host memory (H2D), step, and
            obs/reward/done read back to pinned host memory (D2H) inside the timed region, every step.
  roofline  env_step_kernel: algorithmic bytes (2*N*24 + 2*G + R*2904 + 9 per env-step)
            / measured launch duration vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port (oracle/env_oracle.c, the reference's algorithm restated in C) on the
            host cores, bounded sample.  `--impl reference` times the same port on all host threads as
            its own arm (the reference is pure Python; there is no compiled reference to build).

N > 1 (torchrun): every rank owns its own env batches (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (L, W, people, envs, synthetic layout?)
    "c2": dict(L=36, W=30, people=150, envs=4096, synthetic=False,
               desc="CA-dqn1 single room 36x30, 150 people/env, 4096 envs, env-step only (BASELINE.json configs[1])"),
    "c3": dict(L=256, W=256, people=1000, envs=16384, synthetic=True,
               desc="synthetic Louvre layout 256x256, 1000 people/env, 16384 envs, env-step only (BASELINE.json configs[2] env part)"),
    "c5": dict(L=1024, W=1024, people=20000, envs=512, synthetic=True, exits=8, wall_fill=0.15,
               desc="stress: synthetic 1024x1024 multi-exit museum grid, 20000 people/env, 512 envs per GPU, env-step only (BASELINE.json configs[4] env part)"),
}
L2_BYTES = 126e6


def algorithmic_bytes_per_env_step(L, W, N, R=1):
    """person SoA read+write padded to 24 B, uint8 occupancy read+write, fp32 obs, reward+done."""
    G = (L + 2) * (W + 2)
    return 2 * N * 24 + 2 * G + R * 2904 + 9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def load_traffic(workload):
    """dram bytes per launch from the committed ncu capture of this kernel, if any (profiles/)."""
    p = os.path.join(ROOT, "profiles", "env_step_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region (the profiling recipe's clocks line): an `nvidia-smi -lms 100`
    subprocess started well before the region (its start-up takes longer than a short timed region); `mark()` brackets the
    region and only samples whose timestamp falls inside it are used.  The period stays at 100 ms on purpose: polling at
    20 ms (nvidia-smi) or 5 ms (an in-process NVML thread) slowed the 58 us env-step launches by 11 % (measured)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        import datetime
        if self.t0 is None:
            self.t0 = datetime.datetime.now()
        else:
            self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.1)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows, all_rows = [], []
        with open(self.f.name) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) >= 8:
                    try:
                        ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                        row = (float(c[1]), float(c[2]), float(c[3]), c[4:8])
                    except ValueError:
                        continue
                    all_rows.append(row)
                    if self.t0 and self.t1 and self.t0 <= ts <= self.t1 + datetime.timedelta(milliseconds=100):
                        rows.append(row)
        os.unlink(self.f.name)
        out["samples_in_region"] = len(rows)
        if not rows:                       # region shorter than the sampling period: fall back to the samples under load
            rows = [r for r in all_rows if r[2] > 250.0]
        if rows:
            sm = sorted(r[0] for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = rows[0][1]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = sorted({names[k] for r in rows for k in range(4) if r[3][k].lower().startswith("active")})
            out["samples"] = len(rows)
        return out


def make_layout(wl):
    from dqn_marl_b200.layout import Layout
    if wl["synthetic"]:
        return Layout.synthetic(wl["L"], wl["W"], n_exits=wl.get("exits", 1), wall_fill=wl.get("wall_fill", 0.10), seed=2024)
    return Layout.reference_room(wl["L"], wl["W"])


# ---------------------------------------------------------------------------------------------------
def cpu_port_throughput(layout, wl, n_envs, steps, warm, threads, seed=2026):
    """agent-steps/s of the oracle port on `threads` host threads (same layout, people, reset policy)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from oracle import LayoutTables, OracleBatch
    tabs = LayoutTables.from_layout(layout)
    batch = OracleBatch(tabs, n_envs, wl["people"], 1, seed=seed, threads=threads, reset_robots=1, reset_fire=1)
    batch.reset()
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, 5, size=(steps + warm, n_envs, 1)).astype(np.int32)
    for t in range(warm):
        batch.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warm, warm + steps):
        batch.step(acts[t])
    dt = time.perf_counter() - t0
    return n_envs * wl["people"] * steps / dt, dt


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port on all host threads
    (the reference itself is pure Python and cannot travel to the GPU box; BASELINE.md quotes it at
    ~1.9e4 agent-steps/s/core).  Each step = one pass over the whole batch of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layout = make_layout(wl)
    threads = os.cpu_count() or 1
    # bounded sample: size one step so that K+W steps take about a minute at ~2e6 agent-steps/s/thread
    budget_agent_steps = 60.0 * 2.0e6 * threads
    n_envs = int(budget_agent_steps / max(1, args.steps + args.warmup) / wl["people"])
    n_envs = max(threads * 4, min(wl["envs"], n_envs))
    value, dt = cpu_port_throughput(layout, wl, n_envs, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "env agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "envs_per_step": n_envs, "people": wl["people"]},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_envs} envs x {wl['people']} people x {args.steps} steps, oracle/env_oracle.c, {threads} threads"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from dqn_marl_b200.envs import VecEvacuationEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    layout = make_layout(wl)
    E, N = wl["envs"], wl["people"]
    state_bytes = E * (160 * 0 + ((N + 15) // 16 * 16) * 21 + ((layout.L + 2) * ((layout.W + 2 + 31) // 32) + 3) // 4 * 16 + 96 + 2904 + 13)
    n_rot = max(2, int(L2_BYTES * 1.5 / state_bytes) + 1)
    envs = [VecEvacuationEnv(layout, E, N, device=dev, seed=2026, env_id_base=(rank * n_rot + b) * E,
                             strict_reference=False, auto_reset=True) for b in range(n_rot)]
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n_act = 64
    actions = torch.randint(0, 5, (n_act, E, 1), generator=g, device=dev, dtype=torch.int32)
    obs = [torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev) for _ in range(n_rot)]
    rew = [torch.empty((E,), dtype=torch.float64, device=dev) for _ in range(n_rot)]
    don = [torch.empty((E,), dtype=torch.uint8, device=dev) for _ in range(n_rot)]

    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get('MQ_BENCH_NO_SAMPLER')) else None      # started early: nvidia-smi needs ~100 ms before its first sample
    for b, env in enumerate(envs):
        env.reset()
    # prime: bring every batch to a desynchronised mid-episode mix (untimed)
    for t in range(args.prime):
        for b, env in enumerate(envs):
            env.step_into(actions[t % n_act], obs[b], rew[b], don[b])
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    step_no = 0
    bound = [envs[b].bind_step(obs[b], rew[b], don[b]) for b in range(n_rot)]     # pointer conversions done once
    act_rows = [actions[k] for k in range(n_act)]

    def one_step():
        nonlocal step_no
        bound[step_no % n_rot](act_rows[step_no % n_act])
        step_no += 1

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize(dev)
    launches0 = sum(e.launch_count for e in envs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize(dev)
    if sampler:
        sampler.mark()
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    torch.cuda.synchronize(dev); barrier()
    if sampler:
        sampler.mark()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = sum(e.launch_count for e in envs) - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * E * N * args.steps / (elapsed_ms * 1e-3)

    # L2-resident variant (what a production loop that steps ONE batch sees) — reported, not the headline
    torch.cuda.synchronize(dev)
    ev0.record()
    for k in range(args.steps):
        envs[0].step_into(actions[k % n_act], obs[0], rew[0], don[0])
    ev1.record(); torch.cuda.synchronize(dev)
    warm_ms = ev0.elapsed_time(ev1)

    # ---- e2e: host buffers through the host-facing calls (VecEvacuationEnv.step_async / step_wait): every step copies
    #      its actions from pinned host memory (H2D), runs the step kernel and reads obs / reward / done back to pinned host
    #      memory (D2H), all inside the timed region.  Pipelined: the n_rot independent env batches are in flight on their
    #      own streams, so the PCIe copies of one batch overlap the kernel of another (how an asynchronous vector-env
    #      driver calls it).  e2e_sync = the same with a host synchronisation after every step. ----
    h_act = torch.randint(0, 5, (n_act, E, 1), dtype=torch.int32).pin_memory()
    e2e_steps = max(10 * n_rot, min(args.steps, 300) // n_rot * n_rot)

    def e2e_run(steps, sync_each):
        for k in range(steps):
            b = k % n_rot
            if k >= n_rot and not sync_each:
                envs[b].step_wait()                   # the previous step of this batch has landed on the host
            envs[b].step_async(h_act[k % n_act])
            if sync_each:
                envs[b].step_wait()
        for b in range(n_rot):
            envs[b].step_wait()

    torch.cuda.synchronize(dev)
    e2e_run(2 * n_rot, False)
    e2e_vals = []
    for sync_each in (False, True):
        barrier(); torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_run(e2e_steps, sync_each)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_vals.append(world * E * N * e2e_steps / float(t.item()))
    e2e_value, e2e_sync_value = e2e_vals
    h2d = h_act[0].numel() * 4
    d2h = envs[0].h_obs.numel() * 4 + envs[0].h_reward.numel() * 8 + envs[0].h_done.numel()

    replay = run_replay_leg(args, dev) if args.replay_batch > 0 else None
    learner = learner_fp32 = None
    if args.loop_steps > 0:
        del envs, obs, rew, don
        torch.cuda.empty_cache()
        learner = run_learner_loop(args, wl, layout, dev, rank, world, "bf16")
        torch.cuda.empty_cache()
        learner_fp32 = run_learner_loop(args, wl, layout, dev, rank, world, "fp32") if args.fp32_loop else None

    if rank == 0:
        peaks, peak_src = measured_peaks()
        alg = algorithmic_bytes_per_env_step(layout.L, layout.W, N) * E
        kernel_ms = elapsed_ms / args.steps
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        traffic = load_traffic(args.workload)
        line = {
            "metric": "env agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "envs_per_gpu": E, "people": N, "grid": [layout.L, layout.W],
                       "l2": f"inputs larger than L2: timed loop rotates over {n_rot} independent batches "
                             f"({n_rot * state_bytes / 1e6:.0f} MB of state > 126 MB L2)",
                       "reset_policy": "auto-reset, fresh fire per episode (strict_reference=False)",
                       "prime_steps": args.prime},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "value_sync_each_step": e2e_sync_value,
                    "d2h_GBs_per_gpu": e2e_value / world / (E * N) * d2h / 1e9,
                    "note": f"VecEvacuationEnv.step_async/step_wait: pinned host actions in (H2D), obs+reward+done out (D2H) every "
                            f"step; {n_rot} independent env batches in flight on their own streams (copies overlap kernels); "
                            f"value_sync_each_step = one batch at a time with a host sync per step; d2h_GBs_per_gpu = the PCIe "
                            f"device-to-host rate this e2e value corresponds to (the bound of this leg)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "env_step_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "note": "canonical bytes of SURVEY.md §8d (uint8 occupancy, 24 B/person); this build packs "
                                 "occupancy to 1 bit/cell and rewrites only changed person fields, so it moves fewer bytes"},
            "value_l2_resident": E * N * args.steps / (warm_ms * 1e-3),
        }
        if replay is not None:
            line["replay"] = replay
        if args.loop_steps > 0:
            line["learner"] = learner
            if learner_fp32 is not None:
                line["learner_fp32"] = learner_fp32
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n_cpu_envs = min(E, 1024)
            v0, _ = cpu_port_throughput(layout, wl, n_cpu_envs, 5, 2, threads)
            cpu_steps = int(max(20, min(5000, 12.0 * v0 / (n_cpu_envs * N))))        # ~12 s of CPU work
            v, dt = cpu_port_throughput(layout, wl, n_cpu_envs, cpu_steps, 5, threads)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_cpu_envs} envs x {N} people x {cpu_steps} steps of the same workload, "
                                              f"oracle/env_oracle.c on {threads} threads ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_replay_leg(args, dev):
    """Replay gather/scatter bandwidth (11,634 B per sampled transition, 2 x 5817 B per push)."""
    import torch
    from dqn_marl_b200.replay import ReplayRing
    peaks, _ = measured_peaks()
    cap, B, n_push = 1 << 19, args.replay_batch, 16384          # 6.1 GB ring: far larger than L2
    ring = ReplayRing(cap, device=dev, seed=1)
    s = torch.rand((n_push, 726), device=dev); ns = torch.rand((n_push, 726), device=dev)
    a = torch.zeros(n_push, dtype=torch.int32, device=dev); r = torch.zeros(n_push, dtype=torch.float64, device=dev)
    d = torch.zeros(n_push, dtype=torch.uint8, device=dev)
    for _ in range(cap // n_push):
        ring.push(s, a, r, ns, d)
    out = ring.sample(B)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize(dev); ev0.record()
    for k in range(reps):
        ring.sample(B, out=out)
    ev1.record(); torch.cuda.synchronize(dev)
    ms_s = ev0.elapsed_time(ev1) / reps
    torch.cuda.synchronize(dev); ev0.record()
    for k in range(reps):
        ring.push(s, a, r, ns, d)
    ev1.record(); torch.cuda.synchronize(dev)
    ms_p = ev0.elapsed_time(ev1) / reps
    gbs_s = 11634.0 * B / (ms_s * 1e-3) / 1e9
    gbs_p = 2 * 5817.0 * n_push / (ms_p * 1e-3) / 1e9
    res = {"sample": {"batch": B, "ms": ms_s, "transitions_per_s": B / (ms_s * 1e-3), "achieved_GBs": gbs_s, "frac_of_hbm_peak": gbs_s / peaks["hbm_gbs"]},
           "push": {"n": n_push, "ms": ms_p, "transitions_per_s": n_push / (ms_p * 1e-3), "achieved_GBs": gbs_p, "frac_of_hbm_peak": gbs_p / peaks["hbm_gbs"]},
           "ring_capacity": cap, "kernels": ["replay_sample_kernel", "replay_push_kernel"]}
    del ring, out
    torch.cuda.empty_cache()
    return res


def run_learner_loop(args, wl, layout, dev, rank, world, precision="bf16"):
    """Secondary metric of BASELINE.json: learner transitions/s in the full act -> step -> push -> learn loop
    (train_dqn.py:98-125 batched), fp32 parity path of the Q-network, gradient all-reduce over NCCL when N > 1."""
    import torch
    import torch.distributed as dist
    from dqn_marl_b200.runners.train_dqn_vec import VecTrainer
    E, N, B = wl["envs"], wl["people"], args.learner_batch
    torch.manual_seed(0)
    tr = VecTrainer(layout, E, N, dev, dict(batch_size=B, learning_rate=1e-4, gamma=0.99, epsilon=1.0, epsilon_min=0.02,
                                             epsilon_decay=0.9995, dropout="train", precision=precision),
                    env_id_base=rank * E, seed=2026, replay_capacity=max(1 << 17, 4 * E))
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    seg = [0.0, 0.0, 0.0, 0.0]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.loop_steps):
        a, e = tr.agent, tr.env
        o, o2 = tr.obs[tr.cur], tr.obs[tr.cur ^ 1]
        ev[0].record()
        actions = a.act_batch(o, training=True)
        ev[1].record()
        e.step_into(actions, o2, tr.reward, tr.done)
        ev[2].record()
        a.remember_batch(o, actions, tr.reward, o2, tr.done)
        ev[3].record()
        a.learn_device()
        ev[4].record()
        tr.cur ^= 1
        torch.cuda.synchronize(dev)
        for k in range(4):
            seg[k] += ev[k].elapsed_time(ev[k + 1])
    t1.record(); torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = t0.elapsed_time(t1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    K = args.loop_steps
    launches = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count - l0
    learn_ms = seg[3] / K
    return {"metric": "learner transitions/s (full act->step->push->learn loop)", "value": world * B * K / (total_ms * 1e-3),
            "unit": "transitions/s", "batch_per_gpu": B, "loop_steps": K, "ms_per_loop_step": total_ms / K,
            "env_agent_steps_per_s_in_loop": world * E * N * K / (total_ms * 1e-3),
            "segments_ms": {"act": seg[0] / K, "env_step": seg[1] / K, "replay_push": seg[2] / K, "sample+learn": learn_ms},
            "qnet_dtype": "bf16 tcgen05 tensor cores (conv1-3, fc1, fc2), fp32 accumulate + master weights" if precision == "bf16"
            else "f32 (parity path, CUDA-core FFMA)", "learn_tflops": 155.4e6 * B / (learn_ms * 1e-3) / 1e12,
            "act_tflops": 38.85e6 * E / (seg[0] / K * 1e-3) / 1e12, "gpu_launches": int(launches),
            "allreduce": "nccl all-reduce of the flat 8,157,093-float gradient per learn step" if world > 1 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--prime", type=int, default=150, help="untimed steps per batch before warm-up")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--loop-steps", type=int, default=8, help="steps of the full act/step/push/learn loop (0 = skip)")
    ap.add_argument("--learner-batch", type=int, default=4096)
    ap.add_argument("--fp32-loop", type=int, default=1, help="also run the loop with the fp32 parity path")
    ap.add_argument("--replay-batch", type=int, default=65536, help="replay sample batch of the bandwidth leg (0 = skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
