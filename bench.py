#!/usr/bin/env python3
"""bench.py — headline benchmark of the DQN-MARL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c5]

Metrics (BASELINE.json): env agent-steps/s (headline line) and learner transitions/s (first-class `learner` block of the
same line).  Workload at N=1 = the largest single-GPU configuration, BASELINE.json configs[2] ("C3": synthetic Louvre
layout 256x256, 1000 pedestrians/env, 16384 batched envs, full DQN act/learn loop on 1 B200); `--workload c2` is
configs[1] (36x30 room, 150 people, 4096 envs, env-step only; a short C2 run rides along as `secondary_c2`), `--workload c5`
the stress shape of configs[4].  One env "step" = one fused env-step launch over the whole batch = envs x people agent-steps
(every person counted every step); one learner "step" = act -> env step -> replay push -> sample + learn
(runners/train_dqn.py:98-125 batched) and learns B transitions.

  value     device-resident env throughput: K back-to-back steps, inputs already in HBM, CUDA events.  The timed loop
            ROTATES over independent env batches whose state together exceeds the 126 MB L2 (config.l2), so every step
            reads its state from HBM.
  e2e       same metric through the host-facing call (VecEvacuationEnv.step_async / step_wait): actions from pinned host
            memory (H2D), step, obs/reward/done back to pinned host memory (D2H) inside the timed region, every step.
  roofline  env_step_kernel: algorithmic bytes (2*N*24 + 2*G + R*2904 + 9 per env-step) / measured launch
            duration vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port (oracle/env_oracle.c, the reference's algorithm restated in C) on the host cores, bounded
            sample.
  learner   {value (transitions/s of the full loop), segments, roofline (tensor: 155.4 MFLOP x B / learn ms vs the
            sustained bf16 peak), e2e (the learn step fed from pinned HOST batches, loss read back), cpu_baseline (the torch
            fp32 restatement of DQNAgent.learn(), oracle/qnet_oracle.py, on the host cores at B = 32 and B = batch),
            param_checksum / replicas_identical (data-parallel replicas hold the same weights after the loop)}.
  learner_fp32  the same loop on the 1e-5 parity path (CUDA-core FFMA), fewer steps.
  c1_dropin  BASELINE.json configs[0] (configs/dqn.yaml: one env, 150 people, B = 32) through the drop-in classes, wall clock per
            act / step / remember / learn call; the reference arm reports the same loop on the unmodified reference
            (python_reference.c1_loop).

`--impl reference` times the reference's CPU implementation of the path (the oracle ports; the reference itself is pure
Python and cannot travel to the GPU box) on all host threads, without loading libmarl_b200.so.

N > 1 (torchrun): every rank owns its env batches and replay (weak scaling, no data-path collective); the learner's
gradient is all-reduced over NCCL.
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
    "c2": dict(L=36, W=30, people=150, envs=4096, synthetic=False, learner_batch=4096,
               desc="CA-dqn1 single room 36x30, 150 people/env, 4096 envs, env-step only (BASELINE.json configs[1])"),
    "c3": dict(L=256, W=256, people=1000, envs=16384, synthetic=True, learner_batch=4096,
               desc="synthetic Louvre layout 256x256, 1000 people/env, 16384 envs per GPU, full DQN act/learn loop (BASELINE.json configs[2]; configs[3] when sharded over N GPUs)"),
    "c5": dict(L=1024, W=1024, people=20000, envs=512, synthetic=True, exits=8, wall_fill=0.15, learner_batch=8192,
               desc="stress: synthetic 1024x1024 multi-exit museum grid, 20000 people/env, 512 envs per GPU, replay sample 8192 per GPU (65536 over 8 GPUs) (BASELINE.json configs[4])"),
}
L2_BYTES = 126e6
LEARN_FLOP = 155.4e6        # per learned transition
ACT_FLOP = 38.85e6          # per forward sample


def algorithmic_bytes_per_env_step(L, W, N, R=1):
    """person SoA read+write padded to 24 B, uint8 occupancy read+write, fp32 obs, reward+done."""
    G = (L + 2) * (W + 2)
    return 2 * N * 24 + 2 * G + R * 2904 + 9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback (B200_PROFILING.md)"


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


def make_layout(wl, floor_field=None):
    from dqn_marl_b200.layout import Layout
    if wl["synthetic"]:
        return Layout.synthetic(wl["L"], wl["W"], n_exits=wl.get("exits", 1), wall_fill=wl.get("wall_fill", 0.10), seed=2024,
                                floor_field=floor_field)
    return Layout.reference_room(wl["L"], wl["W"], floor_field=floor_field)


# ---------------------------------------------------------------------------------------------------
# CPU legs (the checker timed as a baseline: the only places bench.py executes oracle/)
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


def cpu_env_baseline(layout, wl, threads, budget_s=12.0):
    N = wl["people"]
    n_cpu_envs = max(threads * 2, min(wl["envs"], max(64, int(150 * 1024 / N))))
    v0, _ = cpu_port_throughput(layout, wl, n_cpu_envs, 3, 1, threads)
    cpu_steps = int(max(6, min(5000, budget_s * v0 / (n_cpu_envs * N))))
    v, dt = cpu_port_throughput(layout, wl, n_cpu_envs, cpu_steps, 2, threads)
    return {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n_cpu_envs} envs x {N} people x {cpu_steps} steps of the same workload, oracle/env_oracle.c on {threads} threads ({dt:.1f} s)"}


def cpu_learner_baseline(batch, budget_s=10.0, seed=0):
    """DQNAgent.learn() (dqn_agent.py:126-168) as restated in torch fp32 (oracle/qnet_oracle.py, pinned to the reference's
    golden numbers by tests/test_agent_ref.py) on the host cores: transitions/s at B = 32 (configs/dqn.yaml) and B = batch."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import qnet_oracle
    threads = torch.get_num_threads()
    out = {}
    for B in (32, batch):
        q, t = qnet_oracle.build_nets(seed)
        opt = torch.optim.Adam(q.parameters(), lr=1e-4)
        g = torch.Generator().manual_seed(seed + B)
        b = (torch.rand((B, 11, 11, 6), generator=g), torch.randint(0, 5, (B,), generator=g), torch.rand((B,), generator=g),
             torch.rand((B, 11, 11, 6), generator=g), torch.rand((B,), generator=g) < 0.05)
        keep = (torch.rand((B, 512), generator=g) >= 0.2)
        qnet_oracle.learn_step(q, t, opt, b, drop_online=keep, drop_target=keep)            # warm-up (allocations, oneDNN primitives)
        n, t0 = 0, time.perf_counter()
        while True:
            qnet_oracle.learn_step(q, t, opt, b, drop_online=keep, drop_target=keep)
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s / 2 or n >= 200:
                break
        out[B] = {"value": B * n / dt, "ms_per_learn": dt / n * 1e3, "steps": n}
    return {"value": out[batch]["value"], "unit": "transitions/s", "cores": threads, "kind": "port",
            "sample": f"torch {torch.__version__} fp32 restatement of DQNAgent.learn() (oracle/qnet_oracle.py) on {threads} host threads: "
                      f"B={batch}: {out[batch]['steps']} learn steps, {out[batch]['ms_per_learn']:.0f} ms each; "
                      f"B=32: {out[32]['steps']} steps, {out[32]['ms_per_learn']:.1f} ms each",
            "b32": {"value": out[32]["value"], "ms_per_learn": out[32]["ms_per_learn"]},
            "batch": batch, "ms_per_learn": out[batch]["ms_per_learn"]}


def python_reference_timing(budget_s=4.0):
    """The UNMODIFIED Python reference (staged under baseline/_ref by scripts/stage_reference.py; git-ignored, travels with the
    snapshot), one process = one core: EvacuationEnv.step at the configs/dqn.yaml room (C1/C2 shape) and at the 256 x 256 / 1000
    people shape, and DQNAgent.learn() at B = 32.  Reported beside the port so that the port's speed-up over the real thing is
    on record; absent tree -> None."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "Louvre_Evacuation", "envs")):
        return None
    import random
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_root)
    try:
        import numpy as np
        import torch
        from Louvre_Evacuation.envs.evacuation_env import EvacuationEnv
        from Louvre_Evacuation.agents.dqn_agent import DQNAgent
    except Exception as e:                                             # noqa: BLE001 — a missing optional import of the reference
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        sys.path.remove(ref_root)
    out = {"kind": "reference", "cores": 1, "source": "baseline/_ref (unmodified Louvre_Evacuation tree)"}
    for key, ctor, people in (("room_36x30_150", lambda: EvacuationEnv(36, 30, None, None, 150), 150),
                              ("grid_256x256_1000", lambda: EvacuationEnv(256, 256, None, [256, 128], 1000), 1000)):
        random.seed(1); np.random.seed(1)
        env = ctor()
        env.reset()
        for _ in range(2):
            env.step(random.randrange(5))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s / 2 and n < 400:
            _, _, done, _ = env.step(random.randrange(5))
            n += 1
            if done:
                env.reset()
        dt = time.perf_counter() - t0
        out[key] = {"agent_steps_per_s": people * n / dt, "ms_per_env_step": dt / n * 1e3, "steps": n}
    torch.manual_seed(0)
    agent = DQNAgent((11, 11, 6), 5, torch.device("cpu"), {"batch_size": 32, "warmup_steps": 0, "memory_size": 2000})
    rng = np.random.default_rng(0)
    for _ in range(64):
        agent.remember(rng.random((11, 11, 6)), int(rng.integers(5)), float(rng.random()), rng.random((11, 11, 6)), False)
    agent.learn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s / 2 and n < 200:
        agent.learn()
        n += 1
    dt = time.perf_counter() - t0
    out["learn_b32"] = {"transitions_per_s": 32 * n / dt, "ms_per_learn": dt / n * 1e3, "steps": n, "torch_threads": torch.get_num_threads()}
    try:
        out["c1_loop"] = c1_loop_timing(EvacuationEnv, DQNAgent, torch.device("cpu"), iters=240, budget_s=max(6.0, budget_s))
    except Exception as e:                                             # noqa: BLE001
        out["c1_loop"] = {"unavailable": f"{type(e).__name__}: {e}"}
    return out


def c1_loop_timing(EvacuationEnv, DQNAgent, device, iters, budget_s, extra_cfg=None):
    """BASELINE.json configs[0] = configs/dqn.yaml: ONE 36x30 env with 150 people, B = 32, replay 50 000, the loop of
    runners/train_dqn.py:98-125 (act -> step -> remember -> learn, reset on done) through the reference's class surface;
    wall clock per call, every call synchronous as the reference API demands.  The same function times the drop-in classes
    (ours arm) and the unmodified reference classes (reference arm, CPU)."""
    import random
    import numpy as np
    random.seed(1); np.random.seed(1)
    cfg = dict(gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995, learning_rate=1e-4, batch_size=32,
               target_update_freq=200, warmup_steps=0, memory_size=50000)
    cfg.update(extra_cfg or {})
    env = EvacuationEnv(width=36, height=30, num_people=150)
    agent = DQNAgent(env.state_size, env.action_size, device, cfg)
    state = env.reset()
    T = dict(act=0.0, step=0.0, remember=0.0, learn=0.0)
    n, warm, t_start = 0, 40, time.perf_counter()
    for it in range(iters):
        t0 = time.perf_counter(); a = agent.act(state, training=True)
        t1 = time.perf_counter(); nstate, r, done, _info = env.step(a)
        t2 = time.perf_counter(); agent.remember(state, a, r, nstate, done)
        t3 = time.perf_counter(); loss = agent.learn() if len(agent.memory) > agent.batch_size else None
        t4 = time.perf_counter()
        state = env.reset() if done else nstate
        if it >= warm and loss is not None:
            T["act"] += t1 - t0; T["step"] += t2 - t1; T["remember"] += t3 - t2; T["learn"] += t4 - t3; n += 1
        if time.perf_counter() - t_start > budget_s and n >= 20:
            break
    if n == 0:
        return None
    ms = {k: v / n * 1e3 for k, v in T.items()}
    it_ms = sum(ms.values())
    return {"workload": "configs/dqn.yaml (BASELINE.json configs[0]): one 36x30 env, 150 people, B = 32, act -> step -> remember -> learn per iteration",
            "ms_per_call": ms, "ms_per_iteration": it_ms, "iterations": n, "agent_steps_per_s": 150 / (it_ms * 1e-3),
            "learned_transitions_per_s": 32 / (it_ms * 1e-3)}


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle ports on all host threads (the reference
    itself is pure Python and cannot travel to the GPU box; BASELINE.md quotes it at ~1.9e4 agent-steps/s/core).  Each step =
    one pass over a bounded batch of the workload.  The layout's floor field is built by the pure-Python restatement
    (oracle/floor_field_py.py): this process never loads libmarl_b200.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from floor_field_py import floor_field
    layout = make_layout(wl, floor_field=floor_field)
    threads = os.cpu_count() or 1
    # bounded sample: size one step so that prime + W + K steps take about a minute at this host's measured speed.  Like the GPU
    # arm, the batch is first run into the steady state of auto-reset (args.prime untimed steps: a step of a fresh, lock-stepped
    # episode costs more than the average step of a long run, on the CPU too), then W warm-up and K timed steps.
    probe_envs = max(threads * 4, min(wl["envs"], 1024))
    v0, _ = cpu_port_throughput(layout, wl, probe_envs, 3, 1, threads)                # agent-steps/s of this host, fresh episodes
    n_envs = int(v0 * 50.0 / max(1, args.prime + args.steps + args.warmup) / wl["people"])      # ~50 s of work in all
    n_envs = max(threads * 4, min(wl["envs"], n_envs))
    value, dt = cpu_port_throughput(layout, wl, n_envs, args.steps, args.prime + args.warmup, threads)
    line = {
        "impl": "reference", "metric": "env agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, wl, layout),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port", "envs_per_step": n_envs,
                         "sample": f"bounded sample of the workload: {n_envs} of its {wl['envs']} envs x {wl['people']} people x {args.steps} steps "
                                   f"after {args.prime} untimed priming steps (steady state of auto-reset, as in the GPU arm), "
                                   f"oracle/env_oracle.c, {threads} threads"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_learner:
        lb = cpu_learner_baseline(args.learner_batch or wl["learner_batch"])
        line["learner"] = {"metric": "learner transitions/s (DQNAgent.learn on the host)", "value": lb["value"], "unit": "transitions/s",
                           "cpu_baseline": lb, "e2e": {"value": lb["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["python_reference"] = python_reference_timing()
    line["loaded_product_library"] = any("libmarl_b200" in ln for ln in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def env_state_bytes(layout, E, N):
    return E * (((N + 15) // 16 * 16) * 21 + ((layout.L + 2) * ((layout.W + 2 + 31) // 32) + 3) // 4 * 16 + 96 + 2904 + 13)


def rotating_batches(layout, E, N):
    """Independent env batches the device-timed loop rotates over so that their state together exceeds L2."""
    return max(2, int(L2_BYTES * 1.5 / env_state_bytes(layout, E, N)) + 1)


def bench_config(args, wl, layout):
    """`config` of the JSON line — the SAME dict in both arms (the reference arm runs a bounded sample of this workload and says
    so in cpu_baseline.sample)."""
    E, N = wl["envs"], wl["people"]
    n_rot, state_bytes = rotating_batches(layout, E, N), env_state_bytes(layout, E, N)
    return {"workload": wl["desc"], "envs_per_gpu": E, "people": N, "grid": [layout.L, layout.W],
            "l2": f"inputs larger than L2: timed loop rotates over {n_rot} independent batches "
                  f"({n_rot * state_bytes / 1e6:.0f} MB of state > 126 MB L2)",
            "reset_policy": "auto-reset, fresh fire per episode (strict_reference=False)",
            "prime_steps": args.prime, "learner_batch_per_gpu": args.learner_batch or wl["learner_batch"]}


def run_env_legs(args, wl, layout, dev, rank, world, sampler=None, e2e=True):
    """Device-resident headline + host-buffer e2e of the env step on one workload -> dict."""
    import torch
    import torch.distributed as dist
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.parallel import max_over_ranks
    E, N = wl["envs"], wl["people"]
    state_bytes = env_state_bytes(layout, E, N)
    # n_rot batches rotate in the device-timed loop (enough to exceed L2); the host-buffer e2e leg keeps n_e2e >= n_rot batches in
    # flight (default 4: with the compact wire form the host expansion of one batch has to hide behind the kernels of the others)
    n_rot = rotating_batches(layout, E, N)
    n_e2e = max(n_rot, int(getattr(args, "batches", 0) or 0) or 4) if e2e else n_rot
    envs = [VecEvacuationEnv(layout, E, N, device=dev, seed=2026, env_id_base=(rank * n_e2e + b) * E,
                             strict_reference=False, auto_reset=True) for b in range(n_e2e)]
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n_act = 64
    actions = torch.randint(0, 5, (n_act, E, 1), generator=g, device=dev, dtype=torch.int32)
    obs = [torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev) for _ in range(n_e2e)]
    rew = [torch.empty((E,), dtype=torch.float64, device=dev) for _ in range(n_e2e)]
    don = [torch.empty((E,), dtype=torch.uint8, device=dev) for _ in range(n_e2e)]
    for env in envs:
        env.reset()
    # prime (untimed): run every batch into its STEADY STATE.  All envs start their first episode together and in lock step
    # (everybody has speed 1.0: steps alternate between "everybody moves" and "nobody moves"), and a step gets cheaper as an
    # episode empties (C3: 0.98 ms at step 0, 0.55 ms at step 600, `scripts/step_time_trace.py`).  After a few episodes of
    # auto-reset the ages are mixed and the step time settles (C3: ~0.80 ms from step ~1500 on): that is what a long-running
    # vector env costs, and it makes the number independent of which K steps are timed.
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

    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    launches = sum(e.launch_count for e in envs) - launches0
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1), dev)
    res = {"value": world * E * N * args.steps / (elapsed_ms * 1e-3), "ms_per_step": elapsed_ms / args.steps, "launches": int(launches),
           "n_rot": n_rot, "state_bytes": state_bytes, "warmup": warm}

    res["value_l2_resident"] = None
    if state_bytes < L2_BYTES:             # what a loop that steps ONE small batch sees — reported, not the headline
        torch.cuda.synchronize(dev)
        ev0.record()
        for k in range(args.steps):
            envs[0].step_into(actions[k % n_act], obs[0], rew[0], don[0])
        ev1.record(); torch.cuda.synchronize(dev)
        res["value_l2_resident"] = E * N * args.steps / (ev0.elapsed_time(ev1) * 1e-3)

    if e2e:
        # host buffers through the host-facing calls: every step copies its actions from pinned host memory (H2D), runs the step
        # kernel and reads obs / reward / done back to pinned host memory (D2H), all inside the timed region.  Pipelined: the n_rot
        # independent env batches are in flight on their own streams, so the PCIe copies of one batch overlap the kernel of
        # another (how an asynchronous vector-env driver calls it).  value_sync_each_step = a host synchronisation per step.
        h_act = torch.randint(0, 5, (n_act, E, 1), dtype=torch.int32).pin_memory()
        per_batch = max(10, min(args.steps // n_rot, 300))
        e2e_steps = per_batch * n_e2e

        def e2e_run(steps, sync_each, wire=False):
            for k in range(steps):
                b = k % n_e2e
                if k >= n_e2e and not sync_each:
                    envs[b].step_wait()                   # the previous step of this batch has landed on the host
                envs[b].step_async(h_act[k % n_act], wire=wire)
                if sync_each:
                    envs[b].step_wait()
            for b in range(n_e2e):
                envs[b].step_wait()

        def timed(sync_each, wire):
            e2e_run(n_e2e, sync_each, wire)
            barrier(); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            e2e_run(e2e_steps, sync_each, wire)
            torch.cuda.synchronize(dev)
            e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
            return world * E * N * e2e_steps / e2e_s

        e2e_run(n_e2e, False)                              # (allocations: streams, pinned buffers, wire buffers, worker threads)
        e2e_run(n_e2e, False, True)
        v_dense, v_sync, v_wire = timed(False, False), timed(True, False), timed(False, True)
        h2d = h_act[0].numel() * 4
        rd = envs[0].h_reward.numel() * 8 + envs[0].h_done.numel()
        d2h = envs[0].h_obs.numel() * 4 + rd
        d2h_wire = envs[0].h_wire.numel() * 4 + rd
        modes = [("dense", v_dense, d2h), ("wire", v_wire, d2h_wire)]
        hybrid = None
        want_hybrid = int(getattr(args, "e2e_hybrid", -1))
        if want_hybrid == 1 or (want_hybrid < 0 and world > 1):
            # HYBRID: on a multi-GPU box both pure modes are bound by the HOST, by different parts of it — dense windows by the
            # DMA / memory path all GPUs share, the wire form by the store bandwidth of this rank's share of the cores.  Sending
            # the fraction f of the envs in wire form and the rest dense uses both at once; f is set so that the two parts take
            # the same time at the rates just measured (identical on every rank: the rates are max-over-ranks values).
            f = round(min(0.95, max(0.05, v_wire / (v_wire + v_dense))), 2)
            v_hyb = timed(False, f)
            n_wire = int(round(f * E))
            d2h_hyb = (E - n_wire) * envs[0].n_robots * 2904 + n_wire * envs[0].n_robots * (envs[0].h_wire.shape[-1] * 4) + rd
            hybrid = {"value": v_hyb, "wire_fraction": f, "d2h_bytes_per_step": d2h_hyb,
                      "note": "step_async(wire=f): the last f of the envs travel in wire form (expanded by the host cores), the others "
                              "as dense windows (written by the GPU's DMA engine); same dense f32 host buffer"}
            modes.append(("hybrid", v_hyb, d2h_hyb))
        mode, best, d2h_best = max(modes, key=lambda m: m[1])
        res["e2e"] = {"value": best, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_best, "mode": mode,
                      "steps": e2e_steps, "value_sync_each_step": v_sync,
                      "d2h_GBs_per_gpu": best / world / (E * N) * d2h_best / 1e9,
                      "dense": {"value": v_dense, "d2h_bytes_per_step": d2h,
                                "note": "step_async(): the dense f32 observation windows cross PCIe (2904 B per window)"},
                      "wire": {"value": v_wire, "d2h_bytes_per_step": d2h_wire, "host_threads": getattr(envs[0], "wire_threads", None),
                               "note": "step_async(wire=True): observations cross PCIe in the compact wire form (544 B per window: channel 2 as "
                                       "f32 + bit planes of channels 1 / 3 / 4) and are expanded to the same dense f32 host buffer by "
                                       "mq_obs_wire_expand inside step_wait(), i.e. inside the timed region"},
                      "hybrid": hybrid,
                      "note": f"VecEvacuationEnv.step_async/step_wait: pinned host actions in (H2D), obs+reward+done out (D2H) every "
                              f"step, the same dense f32 host buffers in every mode; value = the fastest of the transfer modes (named "
                              f"in `mode`), all measured and listed; {n_e2e} independent env batches in flight on their own streams "
                              f"(copies and the host expansion of one overlap the kernels of the others); like the device-timed window, "
                              f"every mode runs on batches in their steady state (see prime_steps); value_sync_each_step = dense mode, "
                              f"one batch at a time with a host sync per step; d2h_GBs_per_gpu = the PCIe device-to-host rate the value "
                              f"corresponds to"}
    for env in envs:
        env.close()
    del envs, obs, rew, don
    torch.cuda.empty_cache()
    return res


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


def run_learner_loop(args, wl, layout, dev, rank, world, precision, loop_steps, overlap=False, host_fed_steps=0):
    """BASELINE.json's second metric: learner transitions/s in the full act -> step -> push -> sample + learn loop
    (train_dqn.py:98-125 batched), gradient all-reduce over NCCL when N > 1.  Per-segment CUDA events are recorded without
    host synchronisation inside the loop and read afterwards."""
    import torch
    import torch.distributed as dist
    from dqn_marl_b200.parallel import max_over_ranks
    from dqn_marl_b200.runners.train_dqn_vec import VecTrainer
    peaks, peak_src = measured_peaks()
    E, N, B = wl["envs"], wl["people"], (args.learner_batch or wl["learner_batch"])
    torch.manual_seed(0)
    tr = VecTrainer(layout, E, N, dev, dict(batch_size=B, learning_rate=1e-4, gamma=0.99, epsilon=1.0, epsilon_min=0.02,
                                             epsilon_decay=0.9995, dropout="train", precision=precision),
                    env_id_base=rank * E, seed=2026, replay_capacity=max(1 << 17, 4 * E), overlap=overlap)
    # the env batch of the loop is run into its steady state first (random actions, untimed; see run_env_legs), so that the env
    # segment of the loop costs what it costs in a long training run and not what the first, lock-stepped episode costs
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + rank)
    rand_act = torch.randint(0, 5, (64, E, 1), generator=g, device=dev, dtype=torch.int32)
    for t in range(args.prime if precision == "bf16" else min(args.prime, 300)):
        tr.env.step_into(rand_act[t % 64], tr.obs[tr.cur], tr.reward, tr.done)
    for _ in range(max(3, -(-B // E) + 2)):
        tr.step()
    torch.cuda.synchronize(dev)
    K = loop_steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)] if not overlap else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    if overlap:
        for _ in range(K):
            tr.step()
        tr.join()
    else:
        a, e = tr.agent, tr.env
        for k in range(K):
            o, o2 = tr.obs[tr.cur], tr.obs[tr.cur ^ 1]
            ev[k][0].record()
            actions = a.act_batch(o, training=True)
            ev[k][1].record()
            e.step_into(actions, o2, tr.reward, tr.done)
            ev[k][2].record()
            a.remember_batch(o, actions, tr.reward, o2, tr.done)
            ev[k][3].record()
            a.learn_device()
            ev[k][4].record()
            tr.cur ^= 1
    t1.record(); torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = max_over_ranks(t0.elapsed_time(t1), dev)
    launches = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count - l0
    res = {"metric": "learner transitions/s (full act->step->push->sample+learn loop)", "value": world * B * K / (total_ms * 1e-3),
           "unit": "transitions/s", "batch_per_gpu": B, "act_batch_per_gpu": E, "loop_steps": K, "ms_per_loop_step": total_ms / K,
           "env_agent_steps_per_s_in_loop": world * E * N * K / (total_ms * 1e-3),
           "qnet_dtype": "bf16 tcgen05 tensor cores (conv1-3, fc1, fc2), fp32 accumulate + master weights" if precision == "bf16"
           else "f32 (parity path, CUDA-core FFMA)", "gpu_launches": int(launches),
           "allreduce": "nccl all-reduce of the flat 8,157,093-float gradient per learn step, overlapped with the conv backward" if world > 1 else None,
           "schedule": "env step + push on a second stream concurrently with sample + learn (learns on the ring as of the previous push)" if overlap
           else "sequential, the reference's order (train_dqn.py:98-125)"}
    if not overlap:
        seg = [sum(ev[k][j].elapsed_time(ev[k][j + 1]) for k in range(K)) / K for j in range(4)]
        learn_ms = seg[3]
        res["segments_ms"] = {"act": seg[0], "env_step": seg[1], "replay_push": seg[2], "sample+learn": learn_ms}
        res["act_tflops"] = ACT_FLOP * E / (seg[0] * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) if precision == "bf16" else None
        tfl = LEARN_FLOP * B / (learn_ms * 1e-3) / 1e12
        res["roofline"] = {"bound": "tensor", "kernel": "sample + learn step (2 forwards, backward, clip + Adam; tcgen05 convs / GEMMs)",
                           "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": (tfl / peak) if peak else None, "traffic": None,
                           "algorithmic_flop_per_step": LEARN_FLOP * B, "ms": learn_ms,
                           "peak_source": peak_src + " bf16_tflops_sustained (kernels timed inside a long step)" if peak else
                           "fp32 CUDA-core path: no tensor-pipe peak applies (B200 fp32 FFMA peak ~74 TFLOP/s)"}
    # data-parallel replicas must hold the same weights after the loop
    chk = tr.agent.net.flat_p.double().sum().reshape(1)
    res["param_checksum"] = float(chk.item())
    if world > 1:
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        res["replicas_identical"] = bool(all(torch.equal(gathered[0], x) for x in gathered))
    if host_fed_steps > 0:
        res["e2e"] = run_host_fed_learner(tr, B, host_fed_steps, dev, world)
    tr.close()
    del tr
    torch.cuda.empty_cache()
    return res


def run_host_fed_learner(tr, B, steps, dev, world):
    """e2e of the learner metric: the learn step fed from HOST memory through the agent's public call — every step copies a
    batch of B transitions (states, actions, rewards, next_states, dones) from pinned host memory (H2D), runs the learn step
    and reads the loss back to the host (D2H), inside the timed region (VecDQNAgent.learn_host: two device slots, the copy of
    batch k+1 overlaps the learn step of batch k)."""
    import torch
    import torch.distributed as dist
    from dqn_marl_b200.parallel import max_over_ranks
    a = tr.agent
    n_host = 4
    g = torch.Generator().manual_seed(5)
    host = [dict(states=torch.rand((B, 11, 11, 6), generator=g).pin_memory(), actions=torch.randint(0, 5, (B,), generator=g).pin_memory(),
                 rewards=torch.rand((B,), generator=g).pin_memory(), next_states=torch.rand((B, 11, 11, 6), generator=g).pin_memory(),
                 dones=(torch.rand((B,), generator=g) < 0.05).to(torch.uint8).pin_memory()) for _ in range(n_host)]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    for k in range(3):
        a.learn_host(host[k % n_host])
    a.learn_host_wait()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(steps):
        a.learn_host(host[k % n_host])
    loss = a.learn_host_wait()
    torch.cuda.synchronize(dev)
    dt = max_over_ranks(time.perf_counter() - t0, dev)
    return {"value": world * B * steps / dt, "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": steps,
            "ms_per_step": dt / steps * 1e3, "last_loss": float(loss),
            "note": "VecDQNAgent.learn_host(batch in pinned host memory): H2D of the batch, learn step, loss read back every step"}


# ---------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

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
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get('MQ_BENCH_NO_SAMPLER')) else None      # started early: nvidia-smi needs ~100 ms before its first sample
    envr = run_env_legs(args, wl, layout, dev, rank, world, sampler)
    clocks = sampler.stop() if sampler else None

    replay = run_replay_leg(args, dev) if args.replay_batch > 0 else None
    learner = learner_fp32 = learner_ovl = None
    if not args.no_learner and args.loop_steps > 0:
        learner = run_learner_loop(args, wl, layout, dev, rank, world, "bf16", args.loop_steps, host_fed_steps=args.host_fed_steps)
        if args.overlap_loop:
            learner_ovl = run_learner_loop(args, wl, layout, dev, rank, world, "bf16", args.loop_steps, overlap=True)
        if args.fp32_loop_steps > 0:
            learner_fp32 = run_learner_loop(args, wl, layout, dev, rank, world, "fp32", args.fp32_loop_steps)
    secondary = None
    if args.workload != "c2" and args.c2_steps > 0 and world == 1:
        wl2 = WORKLOADS["c2"]
        lay2 = make_layout(wl2)
        a2 = argparse.Namespace(**vars(args)); a2.steps = args.c2_steps; a2.warmup = 50
        r2 = run_env_legs(a2, wl2, lay2, dev, rank, world, None, e2e=False)
        peaks, _ = measured_peaks()
        alg2 = algorithmic_bytes_per_env_step(lay2.L, lay2.W, wl2["people"]) * wl2["envs"]
        ach2 = alg2 / (r2["ms_per_step"] * 1e-3) / 1e9
        secondary = {"workload": wl2["desc"], "value": r2["value"], "unit": "agent-steps/s", "ms_per_step": r2["ms_per_step"], "steps": args.c2_steps,
                     "value_l2_resident": r2["value_l2_resident"],
                     "roofline": {"bound": "hbm", "kernel": "env_step_kernel<1,28,false>", "achieved": ach2, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": ach2 / peaks["hbm_gbs"], "traffic": load_traffic("c2"), "algorithmic_bytes_per_launch": alg2}}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        alg = algorithmic_bytes_per_env_step(layout.L, layout.W, N) * E
        kernel_ms = envr["ms_per_step"]
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        n_rot, state_bytes = envr["n_rot"], envr["state_bytes"]
        line = {
            "metric": "env agent-steps/s", "value": envr["value"], "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": envr["warmup"], "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, wl, layout),
            "clocks": clocks,
            "e2e": envr["e2e"],
            "gpu_launches": envr["launches"],
            "roofline": {"bound": "hbm", "kernel": "env_step_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(args.workload),
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "note": "canonical bytes of SURVEY.md §8d (uint8 occupancy, 24 B/person); this build packs "
                                 "occupancy to 1 bit/cell and rewrites only changed person fields, so it moves fewer bytes"},
            "value_l2_resident": envr["value_l2_resident"],
        }
        if replay is not None:
            line["replay"] = replay
        if learner is not None:
            line["learner"] = learner
        if learner_ovl is not None:
            line["learner_overlapped"] = learner_ovl
        if learner_fp32 is not None:
            line["learner_fp32"] = learner_fp32
        if secondary is not None:
            line["secondary_c2"] = secondary
        if world == 1 and args.c1_iters > 0:
            from dqn_marl_b200.envs.evacuation_env import EvacuationEnv
            from dqn_marl_b200.agents.dqn_agent import DQNAgent
            c1 = c1_loop_timing(EvacuationEnv, DQNAgent, dev, iters=args.c1_iters, budget_s=20.0, extra_cfg={"seed": 2})
            if c1 is not None:
                c1["note"] = ("the drop-in classes (dqn_marl_b200.envs.EvacuationEnv / agents.DQNAgent, fp32 parity path) driven exactly like the "
                              "reference's classes; the reference arm's python_reference.c1_loop is the same loop on the unmodified reference (CPU)")
                line["c1_dropin"] = c1
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_env_baseline(layout, wl, threads)
            if learner is not None:
                lb = cpu_learner_baseline(args.learner_batch or wl["learner_batch"])
                learner["cpu_baseline"] = lb
                if learner_fp32 is not None:
                    learner_fp32["cpu_baseline"] = lb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--prime", type=int, default=1500, help="untimed steps per batch before warm-up (default: into the steady state of auto-reset)")
    ap.add_argument("--batches", type=int, default=0, help="independent env batches rotated / in flight (0 = 4, or more if it takes more to exceed L2)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-learner", action="store_true", help="skip the learner legs")
    ap.add_argument("--loop-steps", type=int, default=200, help="steps of the full act/step/push/learn loop, bf16 path (0 = skip)")
    ap.add_argument("--fp32-loop-steps", type=int, default=10, help="steps of the same loop on the fp32 parity path (0 = skip)")
    ap.add_argument("--overlap-loop", type=int, default=1, help="also run the loop with env step + push overlapped with sample + learn")
    ap.add_argument("--host-fed-steps", type=int, default=60, help="learn steps fed from pinned host batches (learner e2e; 0 = skip)")
    ap.add_argument("--learner-batch", type=int, default=0, help="learn batch per GPU (0 = the workload's: 4096, C5: 8192)")
    ap.add_argument("--replay-batch", type=int, default=65536, help="replay sample batch of the bandwidth leg (0 = skip)")
    ap.add_argument("--c2-steps", type=int, default=3000, help="steps of the secondary C2 env-only run (0 = skip)")
    ap.add_argument("--e2e-hybrid", type=int, default=0, help="1 = also time the hybrid dense + wire transfer mode of the e2e leg (measured on 2 GPUs: slower than either pure mode, so off by default)")
    ap.add_argument("--c1-iters", type=int, default=400, help="iterations of the C1 (configs/dqn.yaml) loop through the drop-in classes (0 = skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
