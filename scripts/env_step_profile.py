"""Two consecutive env steps of a bench workload after `prime` untimed ones (a step in which nobody / everybody moves) for an
ncu --set full capture:  python scripts/env_step_profile.py c3 [prime] [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dqn_marl_b200.envs import VecEvacuationEnv

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
prime = int(sys.argv[2]) if len(sys.argv) > 2 else 40
E = int(sys.argv[3]) if len(sys.argv) > 3 else wl["envs"]
dev = torch.device("cuda:0")
env = VecEvacuationEnv(bench.make_layout(wl), E, wl["people"], device=dev, seed=2026, strict_reference=False, auto_reset=True)
obs = torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev)
rew = torch.empty((E,), dtype=torch.float64, device=dev)
don = torch.empty((E,), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1234)
actions = torch.randint(0, 5, (64, E, 1), generator=g, device=dev, dtype=torch.int32)
env.reset()
for t in range(prime + 2):
    env.step_into(actions[t % 64], obs, rew, don)
torch.cuda.synchronize()
print(f"{wl['desc']}: {prime} + 2 steps, {env.launch_count} launches")
