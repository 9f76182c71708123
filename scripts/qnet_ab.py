"""A/B timing of the Q-network's bf16 path under MQ_CONV_EPI8 settings (two epilogue warp sets in the persistent convolutions):
act at B = 16384 and the learn step at B = 4096, CUDA events, one process.   python scripts/qnet_ab.py 0 1 2 3 8 9 15"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200 import _lib
from dqn_marl_b200.agents import qnet_params as qp
from dqn_marl_b200.agents.qnet import QNet

settings = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 4, 8, 15]
d = "cuda:0"
BA, B = 16384, 4096
torch.manual_seed(0)
sd = qp.TorchDQN().state_dict()
obs = (torch.rand((BA, 11, 11, 6), device=d) < 0.3).float()
batch = dict(states=obs[:B].contiguous(), actions=torch.randint(0, 5, (B,), device=d), rewards=torch.randn(B, device=d),
             next_states=obs[B:2 * B].contiguous(), dones=(torch.rand(B, device=d) < 0.1).to(torch.uint8))
hp = _lib.MqHparams()
hp.gamma, hp.lr, hp.beta1, hp.beta2, hp.adam_eps, hp.clip_norm, hp.huber, hp.adam_step = 0.99, 1e-4, 0.9, 0.999, 1e-8, 1.0, 0, 1
for setting in settings:
    os.environ["MQ_CONV_EPI8"] = str(setting)
    net = QNet(d, max_batch=BA)
    net.load_state_dict(sd, "online"); net.sync_target(1.0); net.set_precision("bf16")
    mask = net.dropout_mask(B, 1, 1); mask_a = net.dropout_mask(BA, 1, 2)
    acts = torch.empty((BA,), dtype=torch.int32, device=d)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def learn():
        net.td_backward(batch, hp, mask, mask); net.clip_adam(hp)
    t_act = timed(lambda: net.act(obs, 0.1, 1, 0, 0, 1, mask_a, out=acts), 20)
    t_learn = timed(learn, 20)
    print(f"MQ_CONV_EPI8={setting:2d}: act(B={BA}) {t_act:.3f} ms   learn(B={B}) {t_learn:.3f} ms   loss {net._loss.item():.5f}", flush=True)
    net.close(); del net
    torch.cuda.empty_cache()
