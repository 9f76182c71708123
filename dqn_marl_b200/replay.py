"""ReplayRing — device-resident replay buffer (csrc/replay.cu).

Mirror of the reference's replay: ``DQNAgent.memory = deque(maxlen=memory_size)`` (reference
Louvre_Evacuation/agents/dqn_agent.py:88-89), ``remember()`` (:97-99) and the ``random.sample`` + tensor
stacking at the top of ``learn()`` (:132-140).  ``len(ring)`` is ``len(agent.memory)``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .envs.vec_env import _require_cuda

OBS = _lib.MQ_OBS_SIZE


class ReplayRing:
    def __init__(self, capacity: int, device="cuda", seed: int = 0):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.capacity = int(capacity)
        self.seed = int(seed)
        self.draws = 0
        dev = self.device
        self.state = torch.empty((self.capacity, OBS), dtype=torch.float32, device=dev)
        self.next_state = torch.empty((self.capacity, OBS), dtype=torch.float32, device=dev)
        self.action = torch.zeros((self.capacity,), dtype=torch.int32, device=dev)
        self.reward = torch.zeros((self.capacity,), dtype=torch.float32, device=dev)
        self.done = torch.zeros((self.capacity,), dtype=torch.uint8, device=dev)
        store = _lib.MqReplayStore(*[_lib.ptr(t) for t in (self.state, self.next_state, self.action, self.reward, self.done)])
        h = C.c_void_p()
        _lib.check(self.lib.mq_replay_create(C.byref(h), self.capacity, dev.index, C.byref(store)), "mq_replay_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mq_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.lib.mq_replay_size(self._h))

    @property
    def cursor(self) -> int:
        return int(self.lib.mq_replay_cursor(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mq_replay_launch_count(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- snapshot / restore (the reference checkpoints only the networks, dqn_agent.py:174-182) ------------------------
    def state_dict(self) -> dict:
        """Host copy of the ring: the stored transitions in PHYSICAL slot order (only the filled part), len, write position
        and the sample-draw counter — enough to continue pushing and sampling bit-identically."""
        n = len(self)
        torch.cuda.current_stream(self.device).synchronize()
        return {"capacity": self.capacity, "size": n, "cursor": self.cursor, "draws": self.draws, "seed": self.seed,
                "state": self.state[:n].cpu(), "next_state": self.next_state[:n].cpu(), "action": self.action[:n].cpu(),
                "reward": self.reward[:n].cpu(), "done": self.done[:n].cpu()}

    def load_state_dict(self, sd: dict):
        if int(sd["capacity"]) != self.capacity:
            raise ValueError(f"replay snapshot of capacity {sd['capacity']} does not fit a ring of capacity {self.capacity}")
        n = int(sd["size"])
        for name in ("state", "next_state", "action", "reward", "done"):
            getattr(self, name)[:n].copy_(sd[name])
        _lib.check(self.lib.mq_replay_restore(self._h, n, int(sd["cursor"])), "mq_replay_restore")
        self.draws, self.seed = int(sd["draws"]), int(sd["seed"])

    def push(self, state, action, reward, next_state, done):
        """Append n transitions (device tensors): state/next_state f32 (n, 726)-viewable, action i32 (n,),
        reward f64 (n,) (the env's dtype; cast to f32 on store like dqn_agent.py:138), done u8 (n,)."""
        n = action.numel()
        s = state.reshape(n, OBS)
        ns = next_state.reshape(n, OBS)
        assert s.dtype == torch.float32 and ns.dtype == torch.float32 and s.is_contiguous() and ns.is_contiguous()
        a = action.reshape(n).to(torch.int32)
        r = reward.reshape(n).to(torch.float64)
        d = done.reshape(n).to(torch.uint8)
        _lib.check(self.lib.mq_replay_push(self._h, _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(ns), _lib.ptr(d), n,
                                           self._stream()), "mq_replay_push")

    def sample(self, batch_size: int, inject_idx: Optional[torch.Tensor] = None, out: Optional[dict] = None,
               draw_id: Optional[int] = None, want_idx: bool = False) -> dict:
        """Uniform sample without replacement of the current contents -> dict of device tensors
        states (B,11,11,6) f32, actions (B,) i64, rewards (B,) f32, next_states, dones (B,) u8."""
        B, dev = int(batch_size), self.device
        if out is None:
            out = dict(states=torch.empty((B, 11, 11, 6), dtype=torch.float32, device=dev),
                       actions=torch.empty((B,), dtype=torch.int64, device=dev),
                       rewards=torch.empty((B,), dtype=torch.float32, device=dev),
                       next_states=torch.empty((B, 11, 11, 6), dtype=torch.float32, device=dev),
                       dones=torch.empty((B,), dtype=torch.uint8, device=dev))
        if want_idx and "idx" not in out:
            out["idx"] = torch.empty((B,), dtype=torch.int64, device=dev)
        if inject_idx is not None:
            inject_idx = inject_idx.to(device=dev, dtype=torch.int64).contiguous()
        if draw_id is None:
            draw_id = self.draws
            self.draws += 1
        _lib.check(self.lib.mq_replay_sample(self._h, B, self.seed, int(draw_id), _lib.ptr(inject_idx), _lib.ptr(out["states"]),
                                             _lib.ptr(out["actions"]), _lib.ptr(out["rewards"]), _lib.ptr(out["next_states"]),
                                             _lib.ptr(out["dones"]), _lib.ptr(out.get("idx")), self._stream()), "mq_replay_sample")
        return out
