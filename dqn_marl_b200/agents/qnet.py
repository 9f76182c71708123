"""QNet — device buffers + handle of the CUDA Q-network / learner (csrc/qnet.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from ..envs.vec_env import _require_cuda
from . import qnet_params as qp


class QNet:
    def __init__(self, device="cuda", max_batch: int = 4096, trainable: bool = True):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.max_batch = int(max_batch)
        dev = self.device
        self.flat_p = torch.zeros(qp.TOTAL, dtype=torch.float32, device=dev)     # online parameters
        self.flat_t = torch.zeros(qp.TOTAL, dtype=torch.float32, device=dev)     # target parameters
        self.trainable = trainable
        if trainable:
            self.flat_g = torch.zeros(qp.TOTAL, dtype=torch.float32, device=dev)
            self.flat_m = torch.zeros(qp.TOTAL, dtype=torch.float32, device=dev)
            self.flat_v = torch.zeros(qp.TOTAL, dtype=torch.float32, device=dev)
        bind = _lib.MqQnetBind()
        for k in range(_lib.MQ_QNET_TENSORS):
            off = qp.OFFSETS[k] * 4
            bind.online[k] = self.flat_p.data_ptr() + off
            bind.target[k] = self.flat_t.data_ptr() + off
            if trainable:
                bind.grad[k] = self.flat_g.data_ptr() + off
                bind.adam_m[k] = self.flat_m.data_ptr() + off
                bind.adam_v[k] = self.flat_v.data_ptr() + off
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.mq_qnet_create(C.byref(h), dev.index, self.max_batch, C.byref(bind)), "mq_qnet_create")
        self._h = h
        self.precision = "fp32"
        self._loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._seen = [self.flat_p._version, self.flat_t._version]

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mq_qnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self.lib.mq_qnet_launch_count(self._h))

    # -- parameters ------------------------------------------------------------------------------------
    def load_state_dict(self, sd, which="online"):
        qp.pack(sd, self.flat_p if which == "online" else self.flat_t)
        self.params_changed()

    def params_changed(self):
        """Call after writing flat_p / flat_t directly (the bf16 operand copies are refreshed lazily)."""
        _lib.check(self.lib.mq_qnet_params_changed(self._h), "mq_qnet_params_changed")
        self._seen = [self.flat_p._version, self.flat_t._version]

    def _track_writes(self):
        """Torch-side in-place writes to the parameter buffers or to any view of them (`q_network.parameters()`, a broadcast,
        an external optimizer, a soft update) bump the tensors' version counters: refresh the library's bf16 operand copies
        when that happened since the last call.  The library's own kernels (Adam, target sync) mark their writes themselves."""
        if self.flat_p._version != self._seen[0] or self.flat_t._version != self._seen[1]:
            self.params_changed()

    def set_precision(self, precision: str):
        """'fp32' = CUDA-core parity path, 'bf16' = tcgen05 tensor-core path (conv2/conv3/fc1)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mq_qnet_set_precision(self._h, {"fp32": 0, "bf16": 1}[precision]), "mq_qnet_set_precision")
        self.precision = precision

    def state_dict(self, which="online"):
        return qp.unpack(self.flat_p if which == "online" else self.flat_t)

    # -- compute ---------------------------------------------------------------------------------------
    def forward(self, obs: torch.Tensor, which: str = "online", drop_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """DQNNetwork.forward (dqn_agent.py:35-61): obs (B,11,11,6) f32 on device -> Q (B,5) f32."""
        B = obs.shape[0]
        self._track_writes()
        obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        q = torch.empty((B, 5), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mq_qnet_forward(self._h, 0 if which == "online" else 1, _lib.ptr(obs), B, _lib.ptr(drop_mask),
                                                _lib.ptr(q), self._stream()), "mq_qnet_forward")
        return q

    def act(self, obs: torch.Tensor, eps: float, seed: int, env_id_base: int, tick: int, n_robots: int = 1,
            drop_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, q_out: Optional[torch.Tensor] = None):
        B = obs.shape[0]
        self._track_writes()
        if out is None:
            out = torch.empty((B,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.mq_qnet_act(self._h, _lib.ptr(obs), B, float(eps), int(seed), int(env_id_base), int(tick) & 0xFFFFFFFF,
                                        int(n_robots), _lib.ptr(drop_mask), _lib.ptr(out), _lib.ptr(q_out), self._stream()), "mq_qnet_act")
        return out

    def td_backward(self, batch: dict, hp: "_lib.MqHparams", drop_online=None, drop_target=None, part: int = 0) -> torch.Tensor:
        """part 0 = the whole differentiable half of learn(); 1 / 2 = its two halves (mq_qnet_td_backward_part): after part 1
        the gradients of fc1/fc2/fc3 (flat_g[HEAD_OFFSET:]) are final, part 2 adds the convolution layers."""
        B = batch["actions"].shape[0]
        self._track_writes()
        if part:
            _lib.check(self.lib.mq_qnet_td_backward_part(self._h, _lib.ptr(batch["states"]), _lib.ptr(batch["actions"]), _lib.ptr(batch["rewards"]),
                                                         _lib.ptr(batch["next_states"]), _lib.ptr(batch["dones"]), B, C.byref(hp),
                                                         _lib.ptr(drop_online), _lib.ptr(drop_target), _lib.ptr(self._loss), int(part),
                                                         self._stream()), "mq_qnet_td_backward_part")
            return self._loss
        _lib.check(self.lib.mq_qnet_td_backward(self._h, _lib.ptr(batch["states"]), _lib.ptr(batch["actions"]), _lib.ptr(batch["rewards"]),
                                                _lib.ptr(batch["next_states"]), _lib.ptr(batch["dones"]), B, C.byref(hp),
                                                _lib.ptr(drop_online), _lib.ptr(drop_target), _lib.ptr(self._loss), self._stream()),
                   "mq_qnet_td_backward")
        return self._loss

    def backward(self, obs: torch.Tensor, dq: torch.Tensor, drop_mask: Optional[torch.Tensor] = None):
        """Gradients of the online parameters for an external dL/dQ (B,5) (mq_qnet_backward): fills flat_g."""
        B = obs.shape[0]
        self._track_writes()
        obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        dq = dq.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mq_qnet_backward(self._h, _lib.ptr(obs), _lib.ptr(dq), B, _lib.ptr(drop_mask), self._stream()),
                       "mq_qnet_backward")

    def clip_adam(self, hp: "_lib.MqHparams", grad_scale: float = 1.0) -> torch.Tensor:
        _lib.check(self.lib.mq_qnet_clip_adam(self._h, C.byref(hp), float(grad_scale), _lib.ptr(self._gnorm), self._stream()),
                   "mq_qnet_clip_adam")
        return self._gnorm

    def sync_target(self, tau: float = 1.0):
        _lib.check(self.lib.mq_qnet_sync_target(self._h, float(tau), self._stream()), "mq_qnet_sync_target")

    def dropout_mask(self, B: int, seed: int, counter: int, p: float = 0.2) -> torch.Tensor:
        m = torch.empty((B, 512), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mq_qnet_dropout_mask(_lib.ptr(m), B * 512, float(p), int(seed), int(counter), self._stream()),
                   "mq_qnet_dropout_mask")
        return m
