"""Throw-away profiling aid: cycles per phase of env_step_kernel (lane 0 of every env-warp, summed over envs) for a bench
workload; needs a -DMQ_ENV_TRACE build of env.cu selected through MARL_B200_SO.
    MARL_B200_SO=.../libmarl_b200_envtrace.so python scripts/env_phase_trace.py c2 400"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dqn_marl_b200 import _lib
from dqn_marl_b200.envs import VecEvacuationEnv
lib = _lib.load()
lib.mq_debug_env_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
E = wl["envs"]
dev = torch.device("cuda:0")
env = VecEvacuationEnv(bench.make_layout(wl), E, wl["people"], device=dev, seed=2026, strict_reference=False, auto_reset=True)
obs = torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev)
rew = torch.empty((E,), dtype=torch.float64, device=dev); don = torch.empty((E,), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1234)
actions = torch.randint(0, 5, (64, E, 1), generator=g, device=dev, dtype=torch.int32)
env.reset()
amb_note = True
names = ["stage", "phase 1", "scoring", "proposals", "moves", "reward inputs", "tree+chain", "bitmap+obs", "reward"]
def window(t0, t1, label):
    torch.cuda.synchronize(); lib.mq_debug_env_trace(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(t0, t1): env.step_into(actions[t % 64], obs, rew, don)
    e1.record(); torch.cuda.synchronize()
    out = np.zeros(16, dtype=np.int64); lib.mq_debug_env_trace(ctypes.c_void_p(out.ctypes.data), 0)
    per = out[:9] / (E * (t1 - t0))
    print(f"{label}: {e0.elapsed_time(e1) / (t1 - t0) * 1e3:.1f} us per launch; cycles per env-step by phase: " +
          ", ".join(f"{n} {c:.0f}" for n, c in zip(names, per)) + f"; total {per.sum():.0f}" +
          (f"; movers needing noise draws {out[12]} of {out[13]} ({100.0 * out[12] / max(out[13], 1):.1f} %)" if out[13] else ""))
window(0, 1, "step 0 (nobody moves)"); window(1, 2, "step 1 (everybody moves)"); window(2, 3, "step 2"); window(3, 4, "step 3")
window(4, min(150, steps), "steps 4-149")
if steps > 150: window(150, min(steps, 1500), f"steps 150-{min(steps, 1500) - 1}")
if steps > 1500: window(1500, steps, f"steps 1500-{steps - 1} (steady state of auto-reset)")
