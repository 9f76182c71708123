"""A few steps of the full C3 loop (act -> env step -> replay push -> sample + learn, bf16 path) for the ncu launch list of
the final build:  python scripts/loop_profile.py [workload] [steps] [envs] [learn batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dqn_marl_b200.runners.train_dqn_vec import VecTrainer

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
E = int(sys.argv[3]) if len(sys.argv) > 3 else wl["envs"]
B = int(sys.argv[4]) if len(sys.argv) > 4 else wl["learner_batch"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = VecTrainer(bench.make_layout(wl), E, wl["people"], dev, dict(batch_size=B, epsilon=1.0, dropout="train", precision="bf16"),
                seed=2026, replay_capacity=max(1 << 17, 4 * E))
for _ in range(3):
    tr.step()
torch.cuda.synchronize()
l0 = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    tr.step()
e1.record(); torch.cuda.synchronize()
n = tr.env.launch_count + tr.agent.net.launch_count + tr.agent.memory.launch_count - l0
print(f"{steps} loop steps: {e0.elapsed_time(e1) / steps:.3f} ms per step, {n / steps:.0f} launches per step (library counters)")
