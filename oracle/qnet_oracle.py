"""ORACLE / TEST INFRASTRUCTURE ONLY — plain PyTorch fp32 restatement of the reference's Q-network and learn() math
(Louvre_Evacuation/agents/dqn_agent.py:35-61 forward, :143-160 learn).  Pinned against the golden numbers the real
reference produced (tests/golden/agent_ref.npz) by tests/test_agent_ref.py; the CUDA kernels are then compared with it on
arbitrary batches, and bench.py times it on the host cores as the learner's cpu_baseline (`kind: "port"`).
Never imported by the product package."""
import torch
import torch.nn.functional as F

from dqn_marl_b200.agents.qnet_params import TorchDQN


def build_nets(seed, target_perturb_seed=None, device="cpu"):
    """Weights as DQNAgent.__init__ creates them under torch.manual_seed(seed): q_network, then target_network
    (dqn_agent.py:83-84), target <- online (:95); optionally the deterministic target perturbation of the golden."""
    torch.manual_seed(seed)
    q = TorchDQN()
    t = TorchDQN()
    t.load_state_dict(q.state_dict())
    if target_perturb_seed is not None:
        g = torch.Generator().manual_seed(target_perturb_seed)
        with torch.no_grad():
            for p in t.parameters():
                p.add_(0.01 * torch.randn(p.shape, generator=g))
    return q.to(device), t.to(device)


def forward(net, x, drop_mask=None):
    """dqn_agent.py:35-61; x (B,11,11,6).  drop_mask: (B,512) keep-mask or None (= eval)."""
    x = x.permute(0, 3, 1, 2).contiguous()
    x = F.relu(net.conv1(x)); x = F.relu(net.conv2(x)); x = F.relu(net.conv3(x))
    x = x.reshape(x.size(0), -1)
    x = F.relu(net.fc1(x))
    if drop_mask is not None:
        x = x * drop_mask.to(x.dtype) * (1.0 / (1.0 - 0.2))
    x = F.relu(net.fc2(x))
    return net.fc3(x)


def learn_step(q, t, opt, batch, gamma=0.99, clip=1.0, drop_online=None, drop_target=None, huber=False):
    """dqn_agent.py:143-160.  Returns (loss, total_norm before clipping)."""
    states, actions, rewards, next_states, dones = batch
    cur = forward(q, states, drop_online).gather(1, actions.unsqueeze(1))
    with torch.no_grad():
        nq = forward(t, next_states, drop_target).max(1)[0]
        target = rewards + (gamma * nq * ~dones)
    loss = F.smooth_l1_loss(cur.squeeze(), target) if huber else F.mse_loss(cur.squeeze(), target)
    opt.zero_grad()
    loss.backward()
    total = torch.nn.utils.clip_grad_norm_(q.parameters(), clip)
    opt.step()
    return float(loss), float(total)
