"""Per-step duration of env_step_kernel along an episode (CUDA events around every launch) for a bench workload:
    python scripts/step_time_trace.py c3 300 [n_envs]
Prints the mean launch time per window of 20 steps plus the live/evacuated/dead counts of env 0."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dqn_marl_b200.envs import VecEvacuationEnv

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
E = int(sys.argv[3]) if len(sys.argv) > 3 else wl["envs"]
dev = torch.device("cuda:0")
layout = bench.make_layout(wl)
env = VecEvacuationEnv(layout, E, wl["people"], device=dev, seed=2026, strict_reference=False, auto_reset=True)
obs = torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev)
rew = torch.empty((E,), dtype=torch.float64, device=dev)
don = torch.empty((E,), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1234)
actions = torch.randint(0, 5, (64, E, 1), generator=g, device=dev, dtype=torch.int32)
env.reset()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
evs[0].record()
for t in range(steps):
    env.step_into(actions[t % 64], obs, rew, don)
    evs[t + 1].record()
torch.cuda.synchronize()
ms = [evs[t].elapsed_time(evs[t + 1]) for t in range(steps)]
sc = env.scalars[0].cpu().numpy()
for w in range(0, steps, 20):
    print(f"steps {w:4d}-{min(w + 20, steps) - 1:4d}: {sum(ms[w:w + 20]) / len(ms[w:w + 20]) * 1e3:9.1f} us per launch")
print("env 0 scalars:", sc.tolist())
