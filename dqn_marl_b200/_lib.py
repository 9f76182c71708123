"""ctypes binding of libmarl_b200.so (the C-ABI declared in include/marl_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import of any compute
entry point raises, loudly, with the command that builds it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MARL_B200_SO") or os.path.join(HERE, "libmarl_b200.so")      # override: A/B timing of two builds

MQ_MAX_ROBOTS = 4
MQ_OBS_SIZE = 726
MQ_OBS_WIRE_WORDS = 136
MQ_ENV_SCALARS = 16
MQ_QNET_TENSORS = 12

_vp, _i32, _i64, _u64, _u32, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double


class MqLayout(C.Structure):
    _fields_ = [("L", _i32), ("W", _i32), ("n_fire_steps", _i32),
                ("ctr_box", _i32 * 4), ("int_box", _i32 * 4), ("robot_range", _i32 * 2),
                ("robot_start", (_i32 * 2) * MQ_MAX_ROBOTS), ("reset_obs_center", _i32 * 2), ("obs_exit", _i32 * 2),
                ("dp5", _vp), ("cellinfo", _vp), ("danger_ctr", _vp), ("danger_int", _vp)]


class MqEnvCfg(C.Structure):
    _fields_ = [("n_envs", _i32), ("n_people", _i32), ("n_robots", _i32), ("device", _i32), ("seed", _u64),
                ("env_id_base", _i32), ("max_steps", _i32), ("reset_robots", _i32), ("reset_fire", _i32),
                ("auto_reset", _i32), ("reserved", _i32),
                ("evac_reward", _f64), ("death_penalty", _f64), ("death_acc_penalty", _f64), ("alive_bonus", _f64)]


class MqEnvState(C.Structure):
    _fields_ = [("pos", _vp), ("health", _vp), ("acc", _vp), ("flags", _vp), ("rmap", _vp), ("robots", _vp),
                ("scalars", _vp)]


class MqReplayStore(C.Structure):
    _fields_ = [("state", _vp), ("next_state", _vp), ("action", _vp), ("reward", _vp), ("done", _vp)]


class MqQnetBind(C.Structure):
    _fields_ = [("online", _vp * MQ_QNET_TENSORS), ("target", _vp * MQ_QNET_TENSORS), ("grad", _vp * MQ_QNET_TENSORS),
                ("adam_m", _vp * MQ_QNET_TENSORS), ("adam_v", _vp * MQ_QNET_TENSORS)]


class MqHparams(C.Structure):
    _fields_ = [("gamma", _f32), ("lr", _f32), ("beta1", _f32), ("beta2", _f32), ("adam_eps", _f32),
                ("clip_norm", _f32), ("huber", _i32), ("adam_step", _i32)]


class MarlLibraryMissing(ImportError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/marl_b200.h declares
SIGNATURES = {
    "mq_last_error": (C.c_char_p, []),
    "mq_abi_version": (C.c_int, []),
    "mq_floor_field": (C.c_int, [_i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "mq_floor_field_device": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mq_env_state_sizes": (C.c_int, [C.POINTER(MqEnvCfg), C.POINTER(MqLayout), C.POINTER(_i64), C.POINTER(_i64)]),
    "mq_env_create": (C.c_int, [C.POINTER(_vp), C.POINTER(MqEnvCfg), C.POINTER(MqLayout), C.POINTER(MqEnvState)]),
    "mq_env_create_layouts": (C.c_int, [C.POINTER(_vp), C.POINTER(MqEnvCfg), C.POINTER(MqLayout), _i32, _vp, _i32, C.POINTER(MqEnvState)]),
    "mq_layout_tables_device": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mq_env_destroy": (C.c_int, [_vp]),
    "mq_env_set_reward_coefs": (C.c_int, [_vp, _f64, _f64, _f64, _f64]),
    "mq_env_reset": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mq_env_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mq_env_set_obs_wire": (C.c_int, [_vp, _vp]),
    "mq_obs_wire_expand": (C.c_int, [_vp, _i64, _vp, _i32]),
    "mq_env_unpack_rmap": (C.c_int, [_vp, _vp, _vp]),
    "mq_env_launch_count": (_i64, [_vp]),
    "mq_replay_create": (C.c_int, [C.POINTER(_vp), _i64, _i32, C.POINTER(MqReplayStore)]),
    "mq_replay_destroy": (C.c_int, [_vp]),
    "mq_replay_size": (_i64, [_vp]),
    "mq_replay_cursor": (_i64, [_vp]),
    "mq_replay_launch_count": (_i64, [_vp]),
    "mq_replay_restore": (C.c_int, [_vp, _i64, _i64]),
    "mq_replay_push": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "mq_replay_sample": (C.c_int, [_vp, _i64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mq_qnet_create": (C.c_int, [C.POINTER(_vp), _i32, _i64, C.POINTER(MqQnetBind)]),
    "mq_qnet_destroy": (C.c_int, [_vp]),
    "mq_qnet_forward": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mq_qnet_act": (C.c_int, [_vp, _vp, _i64, _f32, _u64, _u32, _u32, _i32, _vp, _vp, _vp, _vp]),
    "mq_qnet_explore_draw": (C.c_int, [_f32, _u64, _u32, _u32, _u32, C.POINTER(_i32)]),
    "mq_qnet_td_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(MqHparams), _vp, _vp, _vp, _vp]),
    "mq_qnet_td_backward_part": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(MqHparams), _vp, _vp, _vp, _i32, _vp]),
    "mq_qnet_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mq_qnet_clip_adam": (C.c_int, [_vp, C.POINTER(MqHparams), _f32, _vp, _vp]),
    "mq_qnet_sync_target": (C.c_int, [_vp, _f32, _vp]),
    "mq_qnet_dropout_mask": (C.c_int, [_vp, _i64, _f32, _u64, _u64, _vp]),
    "mq_qnet_launch_count": (_i64, [_vp]),
    "mq_qnet_set_precision": (C.c_int, [_vp, _i32]),
    "mq_qnet_params_changed": (C.c_int, [_vp]),
    "mq_gemm_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mq_gemm_bf16_tn": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mq_conv3x3_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "mq_conv3x3_wgrad_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
}


def load():
    """Load libmarl_b200.so or raise MarlLibraryMissing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MarlLibraryMissing(
            f"{SO_PATH} is missing: the CUDA extension has not been built and there is no CPU fallback. "
            "Build it with `python -m dqn_marl_b200.build` (needs nvcc, sm_100a).")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch, also loud
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MqError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mq_last_error()
        raise MqError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def ptr(a):
    """void* of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


def floor_field(L, W, wall, exits, add_term):
    """mq_floor_field -> (L+2, W+2) float64 (reference map.py:127-148)."""
    lib = load()
    wall = np.ascontiguousarray(wall, dtype=np.uint8)
    exits = np.ascontiguousarray(exits, dtype=np.int32).reshape(-1, 2)
    add = np.ascontiguousarray(add_term, dtype=np.float64)
    out = np.empty((L + 2, W + 2), dtype=np.float64)
    check(lib.mq_floor_field(L, W, ptr(wall), ptr(exits), len(exits), ptr(add), ptr(out)), "mq_floor_field")
    return out


def layout_tables_device(L, W, space, barrier, exits, n_exits, obs_exit, stream=None):
    """mq_layout_tables_device: (dp5 (n, L+2, W+2, 8) float64, cellinfo (n, L+2, W+2) uint8) CUDA tensors from the floor fields of
    a batch of layouts (what Layout.build derives on the host for one)."""
    import torch
    lib = load()
    n = space.shape[0]
    space = space.to(torch.float64).contiguous()
    barrier = barrier.to(torch.uint8).contiguous()
    exits = exits.to(torch.int32).contiguous()
    n_exits = n_exits.to(torch.int32).contiguous()
    obs_exit = obs_exit.to(torch.int32).contiguous()
    dp5 = torch.empty((n, L + 2, W + 2, 8), dtype=torch.float64, device=space.device)
    info = torch.empty((n, L + 2, W + 2), dtype=torch.uint8, device=space.device)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(space.device).cuda_stream)
    check(lib.mq_layout_tables_device(L, W, n, ptr(space), ptr(barrier), ptr(exits), exits.shape[1], ptr(n_exits), ptr(obs_exit), ptr(dp5),
                                      ptr(info), st), "mq_layout_tables_device")
    return dp5, info


def floor_field_device(L, W, wall, exits, n_exits, add_term=None, stream=None):
    """mq_floor_field_device: floor fields of a batch of layouts on the GPU (reference map.py:127-148 per layout).
    wall (n, L+2, W+2) uint8 CUDA tensor, exits (n, max_exits, 2) int32, n_exits (n,) int32, add_term (n, L+2, W+2) float64 or None
    -> (space (n, L+2, W+2) float64 CUDA tensor, number of relaxation sweeps)."""
    import torch
    lib = load()
    n = wall.shape[0]
    wall = wall.to(torch.uint8).contiguous()
    exits = exits.to(torch.int32).contiguous()
    n_exits = n_exits.to(torch.int32).contiguous()
    if add_term is not None:
        add_term = add_term.to(torch.float64).contiguous()
    out = torch.empty((n, L + 2, W + 2), dtype=torch.float64, device=wall.device)
    sweeps = C.c_int32(0)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(wall.device).cuda_stream)
    with torch.cuda.device(wall.device):
        check(lib.mq_floor_field_device(L, W, n, ptr(wall), ptr(exits), exits.shape[1], ptr(n_exits), ptr(add_term), ptr(out),
                                        C.cast(C.pointer(sweeps), C.c_void_p), st), "mq_floor_field_device")
    return out, sweeps.value
