"""ORACLE / TEST INFRASTRUCTURE ONLY — golden vectors of the agent path from the UNMODIFIED reference
(Louvre_Evacuation/agents/dqn_agent.py), run here on CPU (torch fp32):  python oracle/make_golden_agent.py

Scenario A (.eval()): Q-values of a fixed batch, two consecutive DQNAgent.learn() calls
            (loss, clipped-gradient norms, parameter checksums after each Adam step), act() greedy choices.
Scenario B (train mode, injected dropout masks): one learn() call with the reference's Dropout(0.2) replaced by a
            mask-applying module fed with masks recorded in the fixture.
Weights are NOT stored (32 MB): they are torch's default init under torch.manual_seed(seed) in the reference's
construction order, which tests/torch_ref.py reproduces with the same torch version (asserted below).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
sys.path.insert(0, HERE); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("MARL_REFERENCE_ROOT", "/root/reference"))
import Louvre_Evacuation.agents.dqn_agent as ref_agent_mod  # noqa: E402
from keyed_draws import sample_indices  # noqa: E402
from util import load_golden, OP_STEP  # noqa: E402

SEED = 20261018
CFG = dict(gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995, learning_rate=1e-4, batch_size=32,
           target_update_freq=200, warmup_steps=0, memory_size=50000)     # configs/dqn.yaml


def transitions(n):
    g = load_golden("traj_room_single.npz")
    out = []
    for f in range(1, len(g["op"])):
        if g["op"][f] == OP_STEP and g["op"][f - 1] in (0, 1):
            out.append((g["obs"][f - 1][0].copy(), int(g["actions"][f][0]) % 5, float(g["reward"][f]), g["obs"][f][0].copy(),
                        bool(g["done"][f])))
        if len(out) == n:
            break
    return out


def checksums(sd):
    cs = {}
    for k, v in sd.items():
        v = v.detach().double().reshape(-1)
        idx = torch.linspace(0, v.numel() - 1, steps=min(16, v.numel())).long()
        cs[k] = dict(sum=float(v.sum()), abssum=float(v.abs().sum()), sample=v[idx].float().tolist(), idx=idx.tolist())
    return cs


class MaskDropout(torch.nn.Module):
    def __init__(self, masks):
        super().__init__()
        self.masks = list(masks)

    def forward(self, x):
        m = self.masks.pop(0)
        return x * m.to(x.dtype) * (1.0 / (1.0 - 0.2))


def build_agent():
    torch.manual_seed(SEED)
    agent = ref_agent_mod.DQNAgent((11, 11, 6), 5, torch.device("cpu"), dict(CFG))
    # the target is a copy of the online net (dqn_agent.py:95); perturb the TARGET deterministically so the two
    # networks differ and the TD target exercises both parameter sets
    g = torch.Generator().manual_seed(SEED + 1)
    with torch.no_grad():
        for p in agent.target_network.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    return agent


def run(scenario, trans, idx_list, masks=None):
    agent = build_agent()
    if scenario == "A":
        agent.q_network.eval(); agent.target_network.eval()
    else:
        agent.q_network.dropout = MaskDropout(masks["online"])
        agent.target_network.dropout = MaskDropout(masks["target"])
    for t in trans:
        agent.remember(*t)
    picks = list(idx_list)

    class _R:                                   # stands in for `random` inside dqn_agent.py (random.sample, :132)
        @staticmethod
        def sample(memory, k):
            idx = picks.pop(0)
            assert len(idx) == k
            return [memory[int(i)] for i in idx]
    saved = ref_agent_mod.random
    ref_agent_mod.random = _R
    out = {}
    try:
        n_learn = 2 if scenario == "A" else 1
        for step in range(n_learn):
            loss = agent.learn()
            out[f"loss{step}"] = float(loss)
            gn = [float(p.grad.norm()) for p in agent.q_network.parameters()]
            out[f"gradnorm{step}"] = gn
            out[f"params{step}"] = checksums(agent.q_network.state_dict())
    finally:
        ref_agent_mod.random = saved
    out["epsilon"] = agent.epsilon
    out["steps"] = agent.steps
    return agent, out


def main():
    from dqn_marl_b200.agents.qnet_params import TorchDQN
    trans = transitions(96)
    B = CFG["batch_size"]
    idx_list = [sample_indices(SEED, k, len(trans), B) for k in range(2)]

    # initial weights are reproducible from the seed by the test-side module
    agent = build_agent()
    torch.manual_seed(SEED)
    mine = TorchDQN()
    for (k, v), (k2, v2) in zip(agent.q_network.state_dict().items(), mine.state_dict().items()):
        assert k == k2 and torch.equal(v, v2), k

    states = torch.tensor(np.array([t[0] for t in trans[:B]]), dtype=torch.float32)
    agent.q_network.eval(); agent.target_network.eval()
    with torch.no_grad():
        q_online = agent.q_network(states).numpy()
        q_target = agent.target_network(states).numpy()
    greedy = [int(agent.act(t[0], training=False)) for t in trans[:B]]
    assert greedy == q_online.argmax(1).tolist()

    _, A = run("A", trans, idx_list)
    gm = torch.Generator().manual_seed(SEED + 2)
    masks = dict(online=[(torch.rand((B, 512), generator=gm) >= 0.2)], target=[(torch.rand((B, 512), generator=gm) >= 0.2)])
    mask_np = dict(online=masks["online"][0].numpy().astype(np.uint8), target=masks["target"][0].numpy().astype(np.uint8))
    _, Bres = run("B", trans, idx_list[:1], masks)

    meta = dict(seed=SEED, cfg=CFG, torch=torch.__version__, numpy=np.__version__, n_transitions=len(trans), A=A, B=Bres,
                greedy=greedy, target_perturb_seed=SEED + 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "agent_ref.npz"),
                        states=np.array([t[0] for t in trans]), actions=np.array([t[1] for t in trans], dtype=np.int64),
                        rewards=np.array([t[2] for t in trans], dtype=np.float64), next_states=np.array([t[3] for t in trans]),
                        dones=np.array([t[4] for t in trans], dtype=np.uint8), idx=np.array(idx_list),
                        q_online=q_online, q_target=q_target, mask_online=mask_np["online"], mask_target=mask_np["target"],
                        meta=np.array(json.dumps(meta)))
    print("agent_ref.npz kB", os.path.getsize(os.path.join(ROOT, "tests", "golden", "agent_ref.npz")) // 1024,
          "loss", A["loss0"], A["loss1"], "B loss", Bres["loss0"])


if __name__ == "__main__":
    main()
