"""Parameter bookkeeping of the Q-network: names/shapes of the reference's ``DQNNetwork.state_dict()``
(reference Louvre_Evacuation/agents/dqn_agent.py:22-31) and the permutations between PyTorch's layouts and the
kernel layouts of csrc/qnet.cu (applied only at the state_dict / checkpoint boundary)."""
from __future__ import annotations

from collections import OrderedDict

import torch

NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
         "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]
TORCH_SHAPES = OrderedDict([
    ("conv1.weight", (32, 6, 3, 3)), ("conv1.bias", (32,)), ("conv2.weight", (64, 32, 3, 3)), ("conv2.bias", (64,)),
    ("conv3.weight", (128, 64, 3, 3)), ("conv3.bias", (128,)), ("fc1.weight", (512, 15488)), ("fc1.bias", (512,)),
    ("fc2.weight", (256, 512)), ("fc2.bias", (256,)), ("fc3.weight", (5, 256)), ("fc3.bias", (5,))])
NUMEL = [int(torch.Size(s).numel()) for s in TORCH_SHAPES.values()]
TOTAL = sum(NUMEL)                       # 8,157,093
OFFSETS = [sum(NUMEL[:k]) for k in range(len(NUMEL))]


def to_internal(name: str, t: torch.Tensor) -> torch.Tensor:
    """PyTorch layout -> kernel layout (flat)."""
    if name.startswith("conv") and name.endswith("weight"):
        return t.permute(2, 3, 1, 0).reshape(-1)                     # [kh][kw][Cin][Cout]
    if name == "fc1.weight":
        return t.reshape(512, 128, 121).permute(0, 2, 1).reshape(-1)  # [n][p][c]
    return t.reshape(-1)


def from_internal(name: str, flat: torch.Tensor) -> torch.Tensor:
    """kernel layout (flat) -> PyTorch layout."""
    shape = TORCH_SHAPES[name]
    if name.startswith("conv") and name.endswith("weight"):
        co, ci = shape[0], shape[1]
        return flat.reshape(3, 3, ci, co).permute(3, 2, 0, 1).contiguous()
    if name == "fc1.weight":
        return flat.reshape(512, 121, 128).permute(0, 2, 1).reshape(512, 15488).contiguous()
    return flat.reshape(shape).clone()


def pack(state_dict, out_flat: torch.Tensor):
    """Write a PyTorch-layout state_dict into a flat kernel-layout buffer."""
    for k, name in enumerate(NAMES):
        src = state_dict[name].detach().to(dtype=torch.float32)
        assert tuple(src.shape) == TORCH_SHAPES[name], f"{name}: shape {tuple(src.shape)} != {TORCH_SHAPES[name]}"
        out_flat[OFFSETS[k]:OFFSETS[k] + NUMEL[k]].copy_(to_internal(name, src).to(out_flat.device))


def unpack(flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((name, from_internal(name, flat[OFFSETS[k]:OFFSETS[k] + NUMEL[k]])) for k, name in enumerate(NAMES))


class TorchDQN(torch.nn.Module):
    """Layer containers only — used for PyTorch's default initialisation (same RNG consumption order as the
    reference's DQNNetwork.__init__, dqn_agent.py:22-31) and for state_dict key/shape compatibility.  No compute
    of the product path goes through this module."""

    def __init__(self):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(6, 32, kernel_size=3, padding=1)
        self.conv2 = torch.nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.conv3 = torch.nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.fc1 = torch.nn.Linear(11 * 11 * 128, 512)
        self.fc2 = torch.nn.Linear(512, 256)
        self.fc3 = torch.nn.Linear(256, 5)
