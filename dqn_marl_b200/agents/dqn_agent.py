"""DQNAgent — drop-in for the reference's agent (reference Louvre_Evacuation/agents/dqn_agent.py:64-191), and
VecDQNAgent, its batched form for thousands of envs per step.

Same constructor / attributes / methods as the reference (``act``, ``remember``, ``learn``,
``update_target_network``, ``save``, ``load``; ``memory``, ``batch_size``, ``epsilon``, ``steps``, ``q_network``,
``target_network``, ``optimizer``), but every tensor lives on the GPU and all arithmetic runs in
libmarl_b200.so: replay ring (csrc/replay.cu) and Q-network / learner kernels (csrc/qnet.cu).

Randomness: epsilon-greedy, replay sampling and dropout use keyed Philox draws; the key comes from
``random.getrandbits(63)`` at construction (so ``random.seed`` makes runs reproducible) unless ``config['seed']``
is given.  ``config['dropout']``: "train" (default — the reference never calls .eval(), so Dropout(0.2) is active in
act() and in both networks during learn(), dqn_agent.py:33,57) or "eval".
"""
from __future__ import annotations

import random
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from .. import _lib, parallel
from ..replay import ReplayRing
from . import qnet_params as qp
from .qnet import QNet


class _QFunction(torch.autograd.Function):
    """Q = q_network(x) as an autograd node: forward = mq_qnet_forward, backward = mq_qnet_backward (the CUDA backward
    kernels of csrc/qnet.cu fed with the incoming dL/dQ).  The parameter gradients land in the agent's flat gradient
    buffer — the tensors `q_network.parameters()` expose as `.grad` — not in autograd's own accumulators; the observation
    gets no gradient (neither does it in the reference: the batch tensors are leaves without requires_grad)."""

    @staticmethod
    def forward(ctx, x, token, view, mask):
        ctx.view, ctx.mask = view, mask
        ctx.save_for_backward(x)
        return view._a.net.forward(x, "online", mask)

    @staticmethod
    def backward(ctx, dq):
        (x,) = ctx.saved_tensors
        ctx.view._a.net.backward(x, dq, ctx.mask)
        return None, None, None, None


class _NetworkView:
    """`agent.q_network` / `agent.target_network`: callable on a (B,11,11,6) tensor, state_dict()/load_state_dict()/
    parameters() like the reference's nn.Module (used by train_dqn.py:75, train_qmix.py:92-100).  Under
    torch.enable_grad() the online network's output carries an autograd node (`_QFunction`), so a loss a runner builds on
    it — train_qmix.py mixes two agents' Q-values first — back-propagates into this agent's gradient buffer with one
    `loss.backward()`; `optimizer.zero_grad()/step()` and `clip_grad_norm_(q_network.parameters(), ..)` then act on it.
    One backward per zero_grad(): the kernels overwrite the gradient buffer instead of accumulating."""

    def __init__(self, agent: "DQNAgent", which: str):
        self._a, self._which = agent, which
        self.training = True
        self._token = None

    def __call__(self, x):
        a = self._a
        if x.dim() == 3:
            x = x.unsqueeze(0)
        mask = a._mask(x.shape[0]) if (self.training and a.dropout_mode == "train") else None
        if self._which == "online" and torch.is_grad_enabled() and a.net.trainable:
            if self._token is None:       # a leaf that requires grad makes autograd call _QFunction.backward
                self._token = torch.zeros((), device=a.net.device, requires_grad=True)
            x = x.to(device=a.net.device, dtype=torch.float32).contiguous()
            return _QFunction.apply(x, self._token, self, mask)
        return a.net.forward(x, self._which, mask)

    forward = __call__

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def state_dict(self):
        return self._a.net.state_dict(self._which)

    def load_state_dict(self, sd):
        self._a.net.load_state_dict(sd, self._which)

    def parameters(self):
        """The 12 parameter tensors as views of the flat buffer (kernel layouts); for the online network `.grad` is the
        matching view of the flat gradient buffer, so torch.nn.utils.clip_grad_norm_ works on them."""
        online = self._which == "online"
        flat = self._a.net.flat_p if online else self._a.net.flat_t
        out = []
        for k in range(len(qp.NAMES)):
            t = flat[qp.OFFSETS[k]:qp.OFFSETS[k] + qp.NUMEL[k]]
            if online and self._a.net.trainable:
                t.grad = self._a.net.flat_g[qp.OFFSETS[k]:qp.OFFSETS[k] + qp.NUMEL[k]]
            out.append(t)
        return out

    def to(self, *_a, **_k):
        return self


class _OptimizerView:
    """`agent.optimizer` with torch.optim.Adam's state_dict format (dqn_agent.py:85,179,189)."""

    def __init__(self, agent: "DQNAgent"):
        self._a = agent

    def zero_grad(self, set_to_none: bool = False):
        self._a.net.flat_g.zero_()

    def step(self):
        a = self._a
        a._adam_t += 1
        hp = a._hparams(clip=float("inf"))
        a.net.clip_adam(hp)

    def state_dict(self):
        a = self._a
        state = {}
        if a._adam_t > 0:
            m, v = qp.unpack(a.net.flat_m), qp.unpack(a.net.flat_v)
            for k, name in enumerate(qp.NAMES):
                state[k] = {"step": torch.tensor(float(a._adam_t)), "exp_avg": m[name].cpu(), "exp_avg_sq": v[name].cpu()}
        group = {"lr": a.learning_rate, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(qp.NAMES)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        a = self._a
        st = sd.get("state", {})
        if st:
            m = OrderedDict((name, st[k]["exp_avg"]) for k, name in enumerate(qp.NAMES))
            v = OrderedDict((name, st[k]["exp_avg_sq"]) for k, name in enumerate(qp.NAMES))
            qp.pack(m, a.net.flat_m)
            qp.pack(v, a.net.flat_v)
            a._adam_t = int(float(st[0]["step"]))
        else:
            a.net.flat_m.zero_(); a.net.flat_v.zero_(); a._adam_t = 0
        groups = sd.get("param_groups")
        if groups:
            a.learning_rate = float(groups[0].get("lr", a.learning_rate))


class DQNAgent:
    def __init__(self, state_size, action_size, device, config, max_batch: Optional[int] = None):
        self.state_size, self.action_size = state_size, action_size
        assert tuple(state_size) == (11, 11, 6) and action_size == 5, "the CUDA Q-network is the reference's 11x11x6 -> 5 net"
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type != "cuda":
            raise RuntimeError(f"dqn_marl_b200.DQNAgent needs a CUDA device (got {device!r}); there is no CPU fallback")
        self.device = dev
        # dqn_agent.py:73-80 (note the default warmup_steps = 1000 when the key is absent)
        self.gamma = config.get("gamma", 0.99)
        self.epsilon = config.get("epsilon", 1.0)
        self.epsilon_min = config.get("epsilon_min", 0.02)
        self.epsilon_decay = config.get("epsilon_decay", 0.9995)
        self.learning_rate = config.get("learning_rate", 0.0001)
        self.batch_size = config.get("batch_size", 32)
        self.target_update_freq = config.get("target_update_freq", 200)     # stored, never used (quirk Q13)
        self.warmup_steps = config.get("warmup_steps", 1000)
        self.huber = bool(config.get("huber", False))
        self.dropout_mode = config.get("dropout", "train")
        self.seed = int(config["seed"]) if "seed" in config else random.getrandbits(63)
        memory_size = config.get("memory_size", 50000)

        self.net = QNet(dev, max_batch=max_batch or max(self.batch_size, 64), trainable=True)
        # 'fp32' (default): CUDA-core path within 1e-5 of the reference; 'bf16': tcgen05 tensor-core path for throughput
        self.precision = config.get("precision", "fp32")
        # PyTorch default init, same RNG consumption order as the reference (q_network, then target_network)
        online, _target = qp.TorchDQN(), qp.TorchDQN()
        self.net.load_state_dict(online.state_dict(), "online")
        self.q_network = _NetworkView(self, "online")
        self.target_network = _NetworkView(self, "target")
        self.optimizer = _OptimizerView(self)
        self.memory = ReplayRing(memory_size, device=self.net.device, seed=self.seed)
        self.steps = 0
        self._adam_t = 0
        self._act_calls = 0
        self._mask_calls = 0
        self.update_target_network()
        if self.precision != "fp32":
            self.net.set_precision(self.precision)

    # ------------------------------------------------------------------
    def _hparams(self, clip: float = 1.0) -> "_lib.MqHparams":
        hp = _lib.MqHparams()
        hp.gamma, hp.lr, hp.beta1, hp.beta2, hp.adam_eps = self.gamma, self.learning_rate, 0.9, 0.999, 1e-8
        hp.clip_norm = clip if np.isfinite(clip) else 3.0e38
        hp.huber, hp.adam_step = int(self.huber), max(self._adam_t, 1)
        return hp

    def _mask(self, B: int):
        self._mask_calls += 1
        return self.net.dropout_mask(B, self.seed, self._mask_calls)

    def _to_dev(self, state) -> torch.Tensor:
        if isinstance(state, np.ndarray):
            state = torch.from_numpy(state.astype(np.float32))      # dqn_agent.py:109
        return state.to(device=self.net.device, dtype=torch.float32)

    # ------------------------------------------------------------------
    # -- single-transition fast path (the reference's loop, train_dqn.py:98-125: one env, numpy states) ----------------------
    _RM_SLOTS = 8
    _RM_BYTES = 2 * 726 * 4 + 8 + 4 + 4          # state f32 | next_state f32 | reward f64 | action i32 | done u8 (+ pad)

    def _staging(self):
        st = getattr(self, "_stage", None)
        if st is None:
            d, nb, ob = self.net.device, self._RM_BYTES, 726 * 4
            hs = torch.zeros((self._RM_SLOTS, nb), dtype=torch.uint8).pin_memory()
            ds = torch.zeros((self._RM_SLOTS, nb), dtype=torch.uint8, device=d)

            def views(buf):
                return [dict(s=b[0:ob].view(torch.float32).view(1, 726), ns=b[ob:2 * ob].view(torch.float32).view(1, 726),
                             r=b[2 * ob:2 * ob + 8].view(torch.float64), a=b[2 * ob + 8:2 * ob + 12].view(torch.int32),
                             d=b[2 * ob + 12:2 * ob + 13]) for b in buf]
            hv = views(hs)
            st = self._stage = dict(hs=hs, ds=ds, hv=[{k: v.numpy() for k, v in e.items()} for e in hv], dv=views(ds), k=0,
                                    ev=[torch.cuda.Event() for _ in range(self._RM_SLOTS)], used=[False] * self._RM_SLOTS,
                                    hx=torch.zeros((1, 726), dtype=torch.float32).pin_memory(),
                                    dx=torch.zeros((1, 726), dtype=torch.float32, device=d),
                                    da=torch.zeros((1,), dtype=torch.int32, device=d), ha=torch.zeros((1,), dtype=torch.int32).pin_memory())
            st["hx_np"], st["ha_np"] = st["hx"].numpy(), st["ha"].numpy()
            import ctypes
            st["explore"] = ctypes.pointer(ctypes.c_int32(0))
        return st

    def remember(self, state, action, reward, next_state, done):
        """dqn_agent.py:97-99.  numpy states (what the reference's runners pass) travel as ONE asynchronous copy of a pinned
        staging slot — state, next_state, reward, action, done — followed by the push kernel; nothing waits for the device."""
        if isinstance(state, np.ndarray) and isinstance(next_state, np.ndarray) and state.size == 726 and next_state.size == 726:
            st = self._staging()
            k = st["k"] % self._RM_SLOTS
            st["k"] += 1
            if st["used"][k]:
                st["ev"][k].synchronize()             # the copy that last read this pinned slot (8 remember() calls ago)
            h, dv = st["hv"][k], st["dv"][k]
            np.copyto(h["s"], state.reshape(1, 726), casting="same_kind")           # float64 -> float32 like dqn_agent.py:136-139
            np.copyto(h["ns"], next_state.reshape(1, 726), casting="same_kind")
            h["r"][0], h["a"][0], h["d"][0] = float(reward), int(action), 1 if done else 0
            with torch.cuda.device(self.net.device):
                st["ds"][k].copy_(st["hs"][k], non_blocking=True)
                st["ev"][k].record()
                st["used"][k] = True
                self.memory.push(dv["s"], dv["a"], dv["r"], dv["ns"], dv["d"])
            return
        d = self.net.device
        self.memory.push(self._to_dev(state).reshape(1, 726).contiguous(), torch.tensor([int(action)], dtype=torch.int32, device=d),
                         torch.tensor([float(reward)], dtype=torch.float64, device=d),
                         self._to_dev(next_state).reshape(1, 726).contiguous(), torch.tensor([1 if done else 0], dtype=torch.uint8, device=d))

    def act(self, state, training=False):
        """dqn_agent.py:101-124 — epsilon-greedy on top of the online forward, one fused launch sequence."""
        if isinstance(state, np.ndarray) and state.size == 726:
            # one state from the host: pinned upload, forward + epsilon-greedy head, pinned read-back, ONE synchronisation
            st = self._staging()
            if training and self.epsilon > 0.0:
                # the exploration draw of this call, evaluated on the host (the same keyed draw the kernel makes): when it says
                # "explore" the reference returns the random action without a forward (dqn_agent.py:103-104) — so do we.  The
                # call and dropout-mask counters advance exactly as on the forward path, so later draws do not depend on it.
                if self.net.lib.mq_qnet_explore_draw(float(self.epsilon), int(self.seed), 0, (self._act_calls + 1) & 0xFFFFFFFF, 0,
                                                     st["explore"]):
                    if self.dropout_mode == "train":
                        self._mask_calls += 1
                    self._act_calls += 1
                    return int(st["explore"].contents.value)
            np.copyto(st["hx_np"], state.reshape(1, 726), casting="same_kind")      # dqn_agent.py:109
            with torch.cuda.device(self.net.device):
                st["dx"].copy_(st["hx"], non_blocking=True)
                mask = self._mask(1) if self.dropout_mode == "train" else None
                self._act_calls += 1
                self.net.act(st["dx"], self.epsilon if training else 0.0, self.seed, 0, self._act_calls, 1, mask, out=st["da"])
                st["ha"].copy_(st["da"], non_blocking=True)
                torch.cuda.current_stream(self.net.device).synchronize()
            return int(st["ha_np"][0])
        x = self._to_dev(state)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        x = x.contiguous()
        B = x.shape[0]
        mask = self._mask(B) if self.dropout_mode == "train" else None
        self._act_calls += 1
        a = self.net.act(x, self.epsilon if training else 0.0, self.seed, 0, self._act_calls, 1, mask)
        return int(a[0].item()) if B == 1 else a

    def learn(self):
        """dqn_agent.py:126-168"""
        if len(self.memory) < self.batch_size or self.steps < self.warmup_steps:
            return None
        B = self.batch_size
        prev = getattr(self, "_learn_batch", None)
        batch = self._learn_batch = self.memory.sample(B, out=prev if prev is not None and prev["actions"].shape[0] == B else None)
        train = self.dropout_mode == "train"
        self._adam_t += 1
        hp = self._hparams()
        loss = self.net.td_backward(batch, hp, self._mask(B) if train else None, self._mask(B) if train else None)
        self._allreduce_grads()
        self.net.clip_adam(hp, self._grad_scale())
        if self.epsilon > self.epsilon_min:
            self.epsilon *= self.epsilon_decay
        self.steps += 1
        return loss.item()

    # hooks for the data-parallel learner (VecDQNAgent overrides)
    def _allreduce_grads(self):
        pass

    def _grad_scale(self) -> float:
        return 1.0

    def update_target_network(self):
        """dqn_agent.py:170-172 — hard copy"""
        self.net.sync_target(1.0)

    def save(self, filepath):
        """dqn_agent.py:174-182 — same dict keys and state_dict layouts as the reference's checkpoints"""
        torch.save({
            "q_network": OrderedDict((k, v.cpu()) for k, v in self.q_network.state_dict().items()),
            "target_network": OrderedDict((k, v.cpu()) for k, v in self.target_network.state_dict().items()),
            "optimizer": self.optimizer.state_dict(),
            "epsilon": self.epsilon,
            "steps": self.steps,
        }, filepath)

    # -- full training state (the reference's checkpoint + what it leaves out) -------------------------
    _COUNTERS = ("steps", "_adam_t", "_act_calls", "_mask_calls", "epsilon", "seed")

    def training_state_dict(self) -> dict:
        """Everything needed to continue training bit-identically: the reference-format checkpoint (networks + Adam state +
        epsilon + steps, dqn_agent.py:176-182) plus the replay ring and the keyed-draw counters the reference does not save."""
        sd = {"checkpoint": {"q_network": OrderedDict((k, v.cpu()) for k, v in self.q_network.state_dict().items()),
                             "target_network": OrderedDict((k, v.cpu()) for k, v in self.target_network.state_dict().items()),
                             "optimizer": self.optimizer.state_dict(), "epsilon": self.epsilon, "steps": self.steps},
              "replay": self.memory.state_dict(),
              "counters": {k: getattr(self, k) for k in self._COUNTERS}}
        for k in ("_tick", "_pushes"):
            if hasattr(self, k):
                sd["counters"][k] = getattr(self, k)
        return sd

    def load_training_state_dict(self, sd: dict):
        ck = sd["checkpoint"]
        self.q_network.load_state_dict(ck["q_network"])
        self.target_network.load_state_dict(ck["target_network"])
        self.optimizer.load_state_dict(ck["optimizer"])
        for k, v in sd["counters"].items():
            setattr(self, k, v)
        self.memory.load_state_dict(sd["replay"])

    def load(self, filepath, allow_pickle: bool = False):
        """dqn_agent.py:184-191.  Reference-format checkpoints hold only tensors, dicts and primitives, so they load with
        `weights_only=True`; `allow_pickle=True` is the explicit opt-in for files that need the unsafe unpickler."""
        ck = torch.load(filepath, map_location="cpu", weights_only=not allow_pickle)
        self.q_network.load_state_dict(ck["q_network"])
        self.target_network.load_state_dict(ck["target_network"])
        self.optimizer.load_state_dict(ck["optimizer"])
        self.epsilon = ck.get("epsilon", self.epsilon_min)
        self.steps = ck.get("steps", 0)


class VecDQNAgent(DQNAgent):
    """Batched agent for VecEvacuationEnv: act on (E*R) observations at once, push E*R transitions per env step,
    learn on large batches; data-parallel over ranks with ONE collective — the all-reduce of the flat gradient."""

    def __init__(self, device, config, n_envs: int, n_robots: int = 1, env_id_base: int = 0, process_group=None):
        cfg = dict(config)
        cfg.setdefault("warmup_steps", 0)
        super().__init__((11, 11, 6), 5, device, cfg, max_batch=max(cfg.get("batch_size", 32), n_envs * n_robots))
        self.n_envs, self.n_robots, self.env_id_base = n_envs, n_robots, env_id_base
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        if self.world > 1:       # identical replicas: rank 0's initial weights everywhere
            torch.distributed.broadcast(self.net.flat_p, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg else 0,
                                        group=self.pg)
            self.net.params_changed()          # flat_p was written behind the library's back: refresh the bf16 operand copies
            self.update_target_network()
        # transitions every rank is guaranteed to have pushed per remember_batch(): env shards may differ by one env
        # (parallel.env_shard), and the learn gate must open on the SAME step everywhere because learn_device() is a collective
        self.min_push = parallel.min_over_ranks(n_envs * n_robots, self.net.device, self.pg)
        self._pushes = 0
        self._actions = torch.zeros((n_envs * n_robots,), dtype=torch.int32, device=self.net.device)
        self._tick = 0

    def act_batch(self, obs: torch.Tensor, training: bool = True) -> torch.Tensor:
        """obs (E, R, 11, 11, 6) f32 on device -> actions (E, R) i32 on device (no host sync)."""
        B = self.n_envs * self.n_robots
        x = obs.reshape(B, 726)
        mask = self._mask(B) if self.dropout_mode == "train" else None
        self.net.act(x, self.epsilon if training else 0.0, self.seed, self.env_id_base, self._tick, self.n_robots, mask, out=self._actions)
        self._tick += 1
        return self._actions.view(self.n_envs, self.n_robots)

    def remember_batch(self, obs, actions, reward, next_obs, done):
        """Push E*R transitions (all device tensors).  The shared env reward/done is repeated per robot
        (train_double_dqn.py:52-56 gives both agents the same scalar reward)."""
        R = self.n_robots
        if R > 1:
            reward = reward.repeat_interleave(R)
            done = done.repeat_interleave(R)
        self.memory.push(obs.reshape(-1, 726), actions.reshape(-1), reward, next_obs.reshape(-1, 726), done)
        self._pushes += 1

    def ready_to_learn(self) -> bool:
        """The gate of train_dqn.py:117-118 / dqn_agent.py:128 (`len(memory) > batch_size`, warm-up) evaluated on quantities
        every rank agrees on — the number of remember_batch() calls times the smallest shard's push size and the replicated
        step counter — so that all ranks enter the collective learn step together (a rank-local len(memory) opens one step
        apart when n_envs % world != 0 and the run hangs in NCCL)."""
        return parallel.learn_gate_open(self._pushes, self.min_push, self.memory.capacity, self.batch_size, self.steps, self.warmup_steps)

    def _allreduce_grads(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.net.flat_g, group=self.pg)      # the one exchange step

    def _grad_scale(self) -> float:
        return 1.0 / self.world

    def learn_device(self) -> torch.Tensor:
        """learn() without the host sync of loss.item(): returns the device loss tensor, or None while the (rank-consistent)
        gate of ready_to_learn() is closed."""
        if not self.ready_to_learn():
            return None
        return self.learn_on(self.sample_batch())

    def sample_batch(self) -> dict:
        """random.sample + stacking (dqn_agent.py:132-140) into this agent's resident batch tensors."""
        self._batch = self.memory.sample(self.batch_size, out=getattr(self, "_batch", None))
        return self._batch

    def learn_on(self, batch: dict) -> torch.Tensor:
        """dqn_agent.py:143-166 on a batch that is already on the device (sampled, or copied from the host by learn_host)."""
        B = batch["actions"].shape[0]
        train = self.dropout_mode == "train"
        self._adam_t += 1
        hp = self._hparams()
        m_on, m_tg = (self._mask(B) if train else None), (self._mask(B) if train else None)
        loss = self.grad_step(batch, hp, m_on, m_tg)
        self.net.clip_adam(hp, self._grad_scale())
        if self.epsilon > self.epsilon_min:
            self.epsilon *= self.epsilon_decay
        self.steps += 1
        return loss

    def grad_step(self, batch: dict, hp, m_on=None, m_tg=None) -> torch.Tensor:
        """Loss + backward of this rank's batch into flat_g and, when data-parallel, the SUM over ranks of flat_g (the one exchange
        step); `_grad_scale()` = 1/world turns it into the gradient of the global batch mean inside clip_adam."""
        if self.world > 1:
            # the all-reduce of the fc gradients (99 % of the bytes) runs on NCCL's stream while the convolution backward is
            # still computing; the small conv slice follows (parallel.allreduce_overlapped)
            loss = parallel.allreduce_overlapped(self.net.flat_g, qp.OFFSETS[6],
                                                 lambda: self.net.td_backward(batch, hp, m_on, m_tg, part=1),
                                                 lambda: self.net.td_backward(batch, hp, m_on, m_tg, part=2), self.pg)
        else:
            loss = self.net.td_backward(batch, hp, m_on, m_tg)
        return loss

    # -- host-fed learner (bench.py's learner e2e; a learner process that receives batches from remote actors) --------------
    def learn_host(self, host_batch: dict) -> None:
        """One learn step on a batch that lives in (pinned) HOST memory: states / next_states f32 (B,11,11,6), actions i64 (B,),
        rewards f32 (B,), dones u8 (B,).  The H2D copy runs on a private copy stream into one of two device slots, so the copy
        of the next batch overlaps this batch's learn step; the loss is copied back to pinned host memory every step
        (`learn_host_wait()` returns the latest)."""
        dev = self.net.device
        main = torch.cuda.current_stream(dev)
        B = host_batch["actions"].shape[0]
        hf = getattr(self, "_hf", None)
        if hf is None or hf["B"] != B:
            def slot():
                return dict(states=torch.empty((B, 11, 11, 6), dtype=torch.float32, device=dev), actions=torch.empty((B,), dtype=torch.int64, device=dev),
                            rewards=torch.empty((B,), dtype=torch.float32, device=dev), next_states=torch.empty((B, 11, 11, 6), dtype=torch.float32, device=dev),
                            dones=torch.empty((B,), dtype=torch.uint8, device=dev))
            hf = self._hf = dict(B=B, k=0, cs=torch.cuda.Stream(device=dev), slots=[slot(), slot()],
                                 free=[torch.cuda.Event(), torch.cuda.Event()], ready=[torch.cuda.Event(), torch.cuda.Event()],
                                 loss=torch.zeros(1, dtype=torch.float32).pin_memory())
        s = hf["k"] % 2
        if hf["k"] >= 2:
            hf["cs"].wait_event(hf["free"][s])            # the learn step that last read this slot has finished
        with torch.cuda.stream(hf["cs"]):
            for key, t in hf["slots"][s].items():
                t.copy_(host_batch[key], non_blocking=True)
            hf["ready"][s].record(hf["cs"])
        main.wait_event(hf["ready"][s])
        loss = self.learn_on(hf["slots"][s])
        hf["free"][s].record(main)
        hf["loss"].copy_(loss, non_blocking=True)
        hf["k"] += 1

    def learn_host_wait(self) -> float:
        torch.cuda.current_stream(self.net.device).synchronize()
        return float(self._hf["loss"][0])
