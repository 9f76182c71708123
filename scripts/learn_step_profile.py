"""A few learn steps at B = 4096 on the bf16 path (for the ncu launch list of the learner kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200 import _lib
from dqn_marl_b200.agents.qnet import QNet
from dqn_marl_b200.agents import qnet_params as qp

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
torch.manual_seed(0)
net = QNet("cuda:0", max_batch=B)
net.load_state_dict(qp.TorchDQN().state_dict(), "online")
net.sync_target(1.0)
net.set_precision(prec)
d = "cuda:0"
batch = dict(states=(torch.rand((B, 11, 11, 6), device=d) < 0.3).float(), actions=torch.randint(0, 5, (B,), device=d),
             rewards=torch.randn(B, device=d), next_states=(torch.rand((B, 11, 11, 6), device=d) < 0.3).float(),
             dones=(torch.rand(B, device=d) < 0.1).to(torch.uint8))
hp = _lib.MqHparams()
hp.gamma, hp.lr, hp.beta1, hp.beta2, hp.adam_eps, hp.clip_norm, hp.huber, hp.adam_step = 0.99, 1e-4, 0.9, 0.999, 1e-8, 1.0, 0, 1
mask = net.dropout_mask(B, 1, 1)
for it in range(3):
    net.td_backward(batch, hp, mask, mask)
    net.clip_adam(hp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(5):
    net.td_backward(batch, hp, mask, mask)
    net.clip_adam(hp)
e1.record(); torch.cuda.synchronize()
print(f"{prec} B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per learn step, loss {net._loss.item():.5f}")
