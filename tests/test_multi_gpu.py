"""Tests that need TWO GPUs on the box (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`):
the data-parallel learner on NCCL (tests/dp_check.py under torchrun) and handles living on a device that is not the current one."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@needs2
def test_data_parallel_gradient_and_replicas_on_nccl():
    """sum of the two half-batch gradients / 2 == gradient of the full batch (through td_backward_part + the overlapped
    all-reduce), and bit-identical replicas after K steps, fp32 and bf16 paths."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("DP_CHECK ")][-1]
    res = json.loads(line[len("DP_CHECK "):])
    assert res["world"] == 2 and res["fp32"]["replicas_identical"] and res["bf16"]["replicas_identical"]
    assert res["fp32"]["grad_rel_err"] <= 2e-5


@needs2
def test_handles_on_a_device_that_is_not_current():
    """An agent / env batch / replay ring built on cuda:1 while cuda:0 is the current device (the reference's
    `DQNAgent(state_size, action_size, torch.device('cuda:1'), cfg)`): every C entry point switches to the handle's device for
    the call, so results equal those of the same objects driven with cuda:1 current."""
    from dqn_marl_b200.agents.dqn_agent import DQNAgent
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    lay = Layout.reference_room()

    def run(set_current):
        torch.cuda.set_device(1 if set_current else 0)
        d1 = torch.device("cuda:1")
        torch.manual_seed(3)
        agent = DQNAgent((11, 11, 6), 5, d1, dict(batch_size=16, warmup_steps=0, memory_size=256, seed=8, epsilon=0.2, dropout="train"))
        env = VecEvacuationEnv(lay, 6, 150, device=d1, seed=2, auto_reset=True, strict_reference=False)
        obs = env.reset()
        rng = np.random.default_rng(0)
        rewards, losses = [], []
        for t in range(24):
            a = agent.act(obs[:, 0], training=True)                 # batch of 6 observations
            obs2, r, d = env.step(a.reshape(6, 1))
            for k in range(6):
                agent.remember(obs[k, 0].cpu().numpy(), int(a[k]), float(r[k]), obs2[k, 0].cpu().numpy(), bool(d[k]))
            obs = obs2.clone()
            rewards.append(r.cpu().numpy().copy())
            l = agent.learn()
            if l is not None:
                losses.append(l)
        assert torch.cuda.current_device() == (1 if set_current else 0)
        return np.stack(rewards), np.array(losses), agent.net.flat_p.cpu(), env.pos.cpu()

    ra, la, pa, xa = run(False)
    rb, lb, pb, xb = run(True)
    torch.cuda.set_device(0)
    assert np.array_equal(ra.view(np.uint64), rb.view(np.uint64)) and np.array_equal(la, lb) and len(la) > 5
    assert torch.equal(pa.view(torch.int32), pb.view(torch.int32)) and torch.equal(xa, xb)
