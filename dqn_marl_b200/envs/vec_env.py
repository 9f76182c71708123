"""VecEvacuationEnv — n_envs independent Louvre_Evacuation instances stepped by one fused CUDA launch.

Host-side mirror of the reference's env surface (reference Louvre_Evacuation/envs/evacuation_env.py:
``reset()`` :61, ``step()`` :122; evacuation_env_multi.py for n_robots = 2), batched: actions in,
(obs, reward, done) out as device tensors.  All arithmetic happens in libmarl_b200.so
(csrc/env.cu); this file only owns the tensors and the handle.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..layout import MAX_ROBOTS, DeviceLayoutBatch, Layout

# scalars layout (include/marl_b200.h)
S_FIRE_STEP, S_CUR_STEP, S_PREV_EVAC, S_PREV_DEAD, S_EPISODE, S_TICK, S_EVAC, S_DEAD, S_RPX, S_RPY = range(10)


def _require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("dqn_marl_b200 runs on CUDA devices only (sm_100a); there is no CPU fallback. "
                           f"Requested device: {device!r}, torch.cuda.is_available()={torch.cuda.is_available()}")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


class VecEvacuationEnv:
    """Batch of environments sharing one ``Layout`` — or, with `layout` a list of Layouts (host tables) / a
    ``DeviceLayoutBatch`` (tables built on the device) and `env_layout` giving every env's layout index, a batch over several
    layouts of one grid size (the reference builds one Map per env instance, evacuation_env.py:42-43).

    strict_reference=True keeps the reference's reset quirks (fire step survives reset — Q6; the
    single-robot env keeps its robot cell and only centres the reset observation at (15,15) — Q7).
    strict_reference=False restarts fire and robots at every reset (what a fresh training run wants).
    """

    def __init__(self, layout: Layout, n_envs: int, num_people: int = 150, device="cuda", seed: int = 0,
                 env_id_base: int = 0, strict_reference: bool = True, auto_reset: bool = True,
                 max_steps: int = 1200, reward_coefs=(50.0, 200.0, 0.5, 1.0), env_layout=None):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        on_device = isinstance(layout, DeviceLayoutBatch)
        layouts = None if on_device else (list(layout) if isinstance(layout, (list, tuple)) else [layout])
        n_layouts = layout.n if on_device else len(layouts)
        if n_layouts > 1:
            if env_layout is None:
                raise ValueError("several layouts need env_layout: the layout index of every env")
            env_layout = np.ascontiguousarray(np.asarray(env_layout).reshape(-1), dtype=np.int32)
            if env_layout.shape[0] != int(n_envs) or env_layout.min() < 0 or env_layout.max() >= n_layouts:
                raise ValueError("env_layout must hold one index in 0..n_layouts-1 per env")
        self.layouts, self.env_layout = layouts, env_layout
        if not on_device:
            layout = layouts[0]
            if any((l.L, l.W, l.n_robots, l.danger_ctr.shape[0]) != (layout.L, layout.W, layout.n_robots, layout.danger_ctr.shape[0]) for l in layouts):
                raise ValueError("the layouts of one batch must share L x W, the number of robots and of fire steps")
        self.layout = layout
        self.n_envs, self.num_people, self.n_robots = int(n_envs), int(num_people), layout.n_robots
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self.strict_reference = bool(strict_reference)
        self._hs = None
        L, W = layout.L, layout.W

        self._keep = []

        def fill(lay, src, obs_exit, tables):
            lay.L, lay.W, lay.n_fire_steps = L, W, src.danger_ctr.shape[0]
            lay.ctr_box[:] = src.ctr_box
            lay.int_box[:] = src.int_box
            lay.robot_range[:] = src.robot_range
            for r in range(MAX_ROBOTS):
                s = src.robot_starts[min(r, len(src.robot_starts) - 1)]
                lay.robot_start[r][0], lay.robot_start[r][1] = int(s[0]), int(s[1])
            lay.reset_obs_center[:] = src.reset_obs_center
            lay.obs_exit[:] = [int(v) for v in obs_exit]
            lay.dp5, lay.cellinfo, lay.danger_ctr, lay.danger_int = tables

        lays = (_lib.MqLayout * n_layouts)()
        if on_device:
            with torch.cuda.device(self.device):
                if layout.device != self.device:
                    raise ValueError(f"the DeviceLayoutBatch lives on {layout.device}, the env batch on {self.device}")
                obs_exits = layout.obs_exit.cpu().numpy()
            G = (L + 2) * (W + 2)
            for k in range(n_layouts):
                fill(lays[k], layout, obs_exits[k], (layout.dp5.data_ptr() + k * G * 64, layout.cellinfo.data_ptr() + k * G,
                                                     layout.danger_ctr.data_ptr(), layout.danger_int.data_ptr()))
            self._keep.append(layout)
        else:
            for k, src in enumerate(layouts):
                arrs = [np.ascontiguousarray(src.dp5, dtype=np.float64), np.ascontiguousarray(src.cellinfo, dtype=np.uint8),
                        np.ascontiguousarray(src.danger_ctr, dtype=np.float64), np.ascontiguousarray(src.danger_int, dtype=np.float64)]
                self._keep.append(arrs)
                fill(lays[k], src, src.obs_exit, [a.ctypes.data for a in arrs])
        lay = lays[0]

        cfg = _lib.MqEnvCfg()
        cfg.n_envs, cfg.n_people, cfg.n_robots = self.n_envs, self.num_people, self.n_robots
        cfg.device = self.device.index
        cfg.seed, cfg.env_id_base, cfg.max_steps = self.seed, int(env_id_base), int(max_steps)
        if strict_reference:
            cfg.reset_robots = 0 if self.n_robots == 1 else 1     # evacuation_env.py:64 vs evacuation_env_multi.py:35
            cfg.reset_fire = 0
        else:
            cfg.reset_robots, cfg.reset_fire = 1, 1
        cfg.auto_reset = int(auto_reset)
        cfg.evac_reward, cfg.death_penalty, cfg.death_acc_penalty, cfg.alive_bonus = [float(v) for v in reward_coefs]

        n_pad, words = C.c_int64(), C.c_int64()
        _lib.check(self.lib.mq_env_state_sizes(C.byref(cfg), C.byref(lay), C.byref(n_pad), C.byref(words)), "mq_env_state_sizes")
        self.n_pad, self.rmap_words = n_pad.value, words.value
        E, dev = self.n_envs, self.device
        with torch.cuda.device(dev):
            self.pos = torch.zeros((E, self.n_pad), dtype=torch.int32, device=dev)
            self.health = torch.zeros((E, self.n_pad), dtype=torch.float64, device=dev)
            self.acc = torch.zeros((E, self.n_pad), dtype=torch.float64, device=dev)
            self.flags = torch.zeros((E, self.n_pad), dtype=torch.uint8, device=dev)
            self.rmap = torch.zeros((E, self.rmap_words), dtype=torch.int32, device=dev)
            self.robots = torch.zeros((E, MAX_ROBOTS, 2), dtype=torch.int32, device=dev)
            self.scalars = torch.zeros((E, _lib.MQ_ENV_SCALARS), dtype=torch.int32, device=dev)
            def starts_of(src):
                return [[int(v) for v in src.robot_starts[min(r, self.n_robots - 1)]] for r in range(MAX_ROBOTS)]
            if layouts is not None and n_layouts > 1:              # every env starts at its own layout's robot cells
                per_layout = torch.tensor([starts_of(src) for src in layouts], dtype=torch.int32, device=dev)
                starts = per_layout[torch.from_numpy(env_layout).to(dev).long()]
            else:
                starts = torch.tensor(starts_of(layout), dtype=torch.int32, device=dev).expand(E, MAX_ROBOTS, 2)
            self.robots[:] = starts                                   # map.py:76-78
            self.scalars[:, S_RPX] = starts[:, 0, 0]
            self.scalars[:, S_RPY] = starts[:, 0, 1]
            self.obs = torch.zeros((E, self.n_robots, 11, 11, 6), dtype=torch.float32, device=dev)
            self.reward = torch.zeros((E,), dtype=torch.float64, device=dev)
            self.done = torch.zeros((E,), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
        st = _lib.MqEnvState(*[_lib.ptr(t) for t in (self.pos, self.health, self.acc, self.flags, self.rmap,
                                                    self.robots, self.scalars)])
        h = C.c_void_p()
        if n_layouts == 1 and not on_device:
            _lib.check(self.lib.mq_env_create(C.byref(h), C.byref(cfg), C.byref(lay), C.byref(st)), "mq_env_create")
        else:
            torch.cuda.synchronize(self.device)            # the device tables must be complete before they are copied
            _lib.check(self.lib.mq_env_create_layouts(C.byref(h), C.byref(cfg), lays, n_layouts, _lib.ptr(env_layout) if n_layouts > 1 else None,
                                                      int(on_device), C.byref(st)), "mq_env_create_layouts")
        self._h = h

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.mq_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_hs", None) is not None:        # an asynchronous host-buffer step may still be in flight on its own stream
            cur.wait_stream(self._hs)
        return C.c_void_p(cur.cuda_stream)

    def set_reward_coefs(self, evac_reward, death_penalty, death_acc_penalty, alive_bonus):
        """EvacuationEnv.EVAC_REWARD etc. are class attributes mutated at runtime (overnight_experiments.py:69-70)."""
        _lib.check(self.lib.mq_env_set_reward_coefs(self._h, evac_reward, death_penalty, death_acc_penalty, alive_bonus),
                   "mq_env_set_reward_coefs")

    def reset(self, env_mask: Optional[torch.Tensor] = None, inject_spawn: Optional[torch.Tensor] = None,
              obs64: Optional[torch.Tensor] = None, obs_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """EvacuationEnv.reset (evacuation_env.py:61-82) for all (or masked) envs -> obs (E, R, 11, 11, 6) f32.
        obs_out: caller-owned f32 tensor of that shape to receive the observations (rows of envs outside the mask are left
        untouched) instead of the internal buffer."""
        if env_mask is not None:
            env_mask = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if inject_spawn is not None:
            inject_spawn = inject_spawn.to(device=self.device, dtype=torch.int16).contiguous()
            assert inject_spawn.shape == (self.n_envs, self.num_people, 2)
        with torch.cuda.device(self.device):
            out = self.obs if obs_out is None else obs_out
            assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == self.obs.numel()
            _lib.check(self.lib.mq_env_reset(self._h, _lib.ptr(env_mask), _lib.ptr(inject_spawn), _lib.ptr(out),
                                             _lib.ptr(obs64), self._stream()), "mq_env_reset")
        return out

    def step(self, actions: torch.Tensor, obs64: Optional[torch.Tensor] = None):
        """EvacuationEnv.step (evacuation_env.py:122-172) for every env.  actions: int (E, R) or (E,) on device.
        Returns (obs f32 (E,R,11,11,6), reward f64 (E,), done u8 (E,)) — views of internal buffers that the next
        call overwrites."""
        a = actions.to(device=self.device, dtype=torch.int32).reshape(self.n_envs, self.n_robots).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mq_env_step(self._h, _lib.ptr(a), _lib.ptr(self.obs), _lib.ptr(obs64), _lib.ptr(self.reward),
                                            _lib.ptr(self.done), self._stream()), "mq_env_step")
        return self.obs, self.reward, self.done

    def step_into(self, actions_i32: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        """Low-overhead variant for benchmarks / trainers: caller-owned, correctly typed device tensors."""
        _lib.check(self.lib.mq_env_step(self._h, _lib.ptr(actions_i32), _lib.ptr(obs), None, _lib.ptr(reward), _lib.ptr(done),
                                        self._stream()), "mq_env_step")

    def bind_step(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor, stream: Optional[torch.cuda.Stream] = None):
        """Pre-bound step for launch-rate-critical loops: returns fn(actions_i32) that enqueues one step into the given
        output tensors on `stream` (default: the current stream at bind time) with the pointer conversions done once —
        a 58 us kernel leaves ~15 us of Python per launch."""
        fn, h = self.lib.mq_env_step, self._h
        po, pr, pd = _lib.ptr(obs), _lib.ptr(reward), _lib.ptr(done)
        st = C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)
        keep = (obs, reward, done)

        def step(actions_i32: torch.Tensor, _keep=keep):
            rc = fn(h, C.c_void_p(actions_i32.data_ptr()), po, None, pr, pd, st)
            if rc:
                _lib.check(rc, "mq_env_step")
        return step

    # ---- host-buffer interface (gym-style step_async / step_wait) ------------------------------------------------
    def step_async(self, actions_host: torch.Tensor, wire=False):
        """Enqueue one step driven from HOST memory on this env batch's own stream: H2D of the (pinned) int32 actions,
        the fused step kernel, D2H of obs / reward / done into pinned host buffers.  Returns immediately; several env
        batches can be in flight so that the PCIe copies of one overlap the kernel of another.  `step_wait()` returns the
        host tensors.  (The reference's step() is synchronous and host-side: evacuation_env.py:122-172.)

        wire=True: the observations travel in the compact wire form (mq_env_set_obs_wire: 544 B per window instead of
        2904 B — channel 2 as f32, channels 1 / 3 / 4 as bit planes) and `step_wait()` expands them on the host
        (mq_obs_wire_expand) into the same pinned f32 buffer, bit-identical to the dense copy.
        wire=f (0 < f < 1): HYBRID — the last round(f * n_envs) envs travel in wire form, the others dense, so that the PCIe
        link (dense part) and the host cores (expansion of the wire part) work side by side; same host buffers."""
        if getattr(self, "_hs", None) is None:
            self._hs = torch.cuda.Stream(device=self.device)
            self._hev = torch.cuda.Event()
            self._d_act = torch.empty((self.n_envs, self.n_robots), dtype=torch.int32, device=self.device)
            self.h_obs = torch.empty(self.obs.shape, dtype=torch.float32).pin_memory()
            self.h_reward = torch.empty((self.n_envs,), dtype=torch.float64).pin_memory()
            self.h_done = torch.empty((self.n_envs,), dtype=torch.uint8).pin_memory()
            self._wire_mode = False
        frac = 1.0 if wire is True else (0.0 if wire is False or wire is None else float(wire))
        n_wire = max(0, min(self.n_envs, int(round(frac * self.n_envs))))
        n_dense = self.n_envs - n_wire
        if n_wire and getattr(self, "d_wire", None) is None:
            self.d_wire = torch.zeros((self.n_envs, self.n_robots, _lib.MQ_OBS_WIRE_WORDS), dtype=torch.int32, device=self.device)
            self.h_wire = torch.zeros(self.d_wire.shape, dtype=torch.int32).pin_memory()
        if bool(n_wire) != self._wire_mode:
            _lib.check(self.lib.mq_env_set_obs_wire(self._h, _lib.ptr(self.d_wire) if n_wire else None), "mq_env_set_obs_wire")
            self._wire_mode = bool(n_wire)
        self._wire_pending = n_wire
        # everything the caller has enqueued so far on its own stream (reset(), step(), writes to the state tensors such as the
        # facade's robot_position setter) happens before this step
        self._hs.wait_stream(torch.cuda.current_stream(self.device))
        a = actions_host.reshape(self.n_envs, self.n_robots)
        assert a.dtype == torch.int32 and a.device.type == "cpu"
        with torch.cuda.stream(self._hs):
            self._d_act.copy_(a, non_blocking=True)
            _lib.check(self.lib.mq_env_step(self._h, _lib.ptr(self._d_act), _lib.ptr(self.obs) if n_dense else None, None, _lib.ptr(self.reward),
                                            _lib.ptr(self.done), C.c_void_p(self._hs.cuda_stream)), "mq_env_step")
            if n_dense:
                self.h_obs[:n_dense].copy_(self.obs[:n_dense], non_blocking=True)
            if n_wire:
                self.h_wire[n_dense:].copy_(self.d_wire[n_dense:], non_blocking=True)
            self.h_reward.copy_(self.reward, non_blocking=True)
            self.h_done.copy_(self.done, non_blocking=True)
            self._hev.record(self._hs)

    def step_wait(self):
        """Block until the step enqueued by step_async() has landed in host memory -> (obs, reward, done) pinned host tensors."""
        self._hev.synchronize()
        n_wire = int(getattr(self, "_wire_pending", 0) or 0)
        if n_wire:
            # host threads of the expansion: this process's share of the cores (torchrun starts LOCAL_WORLD_SIZE of us)
            nthr = getattr(self, "wire_threads", None)
            if nthr is None:
                cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
                nthr = self.wire_threads = max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
            n_dense = self.n_envs - n_wire
            _lib.check(self.lib.mq_obs_wire_expand(C.c_void_p(self.h_wire[n_dense:].data_ptr()), n_wire * self.n_robots,
                                                   C.c_void_p(self.h_obs[n_dense:].data_ptr()), int(nthr)), "mq_obs_wire_expand")
            self._wire_pending = 0
        return self.h_obs, self.h_reward, self.h_done

    def rmap_bytes(self) -> torch.Tensor:
        """People.rmap of every env as uint8 (E, L+2, W+2)."""
        out = torch.empty((self.n_envs, self.layout.L + 2, self.layout.W + 2), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mq_env_unpack_rmap(self._h, _lib.ptr(out), self._stream()), "mq_env_unpack_rmap")
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.mq_env_launch_count(self._h))

    # -- snapshot / restore of the whole batch (the reference has no env checkpoint) ----------------
    _STATE_TENSORS = ("pos", "health", "acc", "flags", "rmap", "robots", "scalars")

    def state_dict(self) -> dict:
        """Host copy of every env's device state.  The keyed draws are functions of (seed, env id, tick / episode counters in
        `scalars`), so these seven tensors plus the constructor arguments continue a run bit-identically."""
        torch.cuda.current_stream(self.device).synchronize()
        if getattr(self, "_hs", None) is not None:
            self._hs.synchronize()
        sd = {k: getattr(self, k).cpu() for k in self._STATE_TENSORS}
        sd["meta"] = {"n_envs": self.n_envs, "num_people": self.num_people, "n_robots": self.n_robots, "L": self.layout.L,
                      "W": self.layout.W, "seed": self.seed, "env_id_base": self.env_id_base, "strict_reference": self.strict_reference}
        return sd

    def load_state_dict(self, sd: dict):
        m = sd["meta"]
        mine = {"n_envs": self.n_envs, "num_people": self.num_people, "n_robots": self.n_robots, "L": self.layout.L, "W": self.layout.W}
        for k, v in mine.items():
            if int(m[k]) != v:
                raise ValueError(f"env snapshot has {k}={m[k]}, this batch has {k}={v}")
        if int(m["seed"]) != self.seed or int(m["env_id_base"]) != self.env_id_base:
            raise ValueError("env snapshot was taken with another seed / env_id_base: the keyed draws would differ")
        for k in self._STATE_TENSORS:
            getattr(self, k).copy_(sd[k])

    # ------------------------------------------------------------------
    def snapshot(self, env: int = 0) -> dict:
        """Host copy of one env's state in the layout of the golden frames (tests / single-env facade)."""
        N = self.num_people
        pos = self.pos[env, :N].cpu().numpy().view(np.uint32)
        sc = self.scalars[env].cpu().numpy()
        return dict(px=(pos & 0xFFFF).astype(np.int16), py=(pos >> 16).astype(np.int16),
                    health=self.health[env, :N].cpu().numpy(), acc=self.acc[env, :N].cpu().numpy(),
                    flags=self.flags[env, :N].cpu().numpy(), rmap=self.rmap_bytes()[env].cpu().numpy(),
                    robots=self.robots[env, :self.n_robots].cpu().numpy().astype(np.int16),
                    fire_step=np.int32(sc[S_FIRE_STEP]), cur_step=np.int32(sc[S_CUR_STEP]), scalars=sc)
