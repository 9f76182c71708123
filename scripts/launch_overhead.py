"""Host-side cost of one env-step launch (Python + ctypes + cudaLaunchKernel) vs its device time, C2 shape.

Reading the output: the "host" figure of the env-step loops is NOT the launch cost — 3000 launches of a ~50 us kernel
fill the 1024-entry launch queue, after which every launch call blocks until a slot frees, so the loop returns after
(3000 - 1024) kernel times.  The true cost of a launch through ctypes is the 3-4 us of the tiny-kernel loops below: the
device-resident bench loop is never host-bound."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dqn_marl_b200.envs import VecEvacuationEnv
wl = bench.WORKLOADS["c2"]; E = wl["envs"]; dev = torch.device("cuda:0")
layout = bench.make_layout(wl)
env = VecEvacuationEnv(layout, E, wl["people"], device=dev, seed=1, strict_reference=False, auto_reset=True)
obs = torch.empty((E, 1, 11, 11, 6), dtype=torch.float32, device=dev); rew = torch.empty((E,), dtype=torch.float64, device=dev); don = torch.empty((E,), dtype=torch.uint8, device=dev)
actions = torch.randint(0, 5, (64, E, 1), device=dev, dtype=torch.int32); rows = [actions[k] for k in range(64)]
env.reset()
fn = env.bind_step(obs, rew, don)
for name, call in (("step_into", lambda k: env.step_into(rows[k % 64], obs, rew, don)), ("bind_step", lambda k: fn(rows[k % 64]))):
    for k in range(200): call(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for k in range(3000): call(k)
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: host {t_host / 3000 * 1e6:.1f} us per launch (loop returned), device {e0.elapsed_time(e1) / 3000 * 1e3:.1f} us per launch")
x = torch.zeros(1024, device=dev)
for k in range(200): x.add_(1.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(3000): x.add_(1.0)
t = time.perf_counter() - t0; torch.cuda.synchronize()
print(f"torch x.add_(1): host {t / 3000 * 1e6:.1f} us per launch")
from dqn_marl_b200 import _lib
import ctypes as C
lib = _lib.load()
m = torch.empty((64, 512), dtype=torch.uint8, device=dev); st = C.c_void_p(torch.cuda.current_stream().cuda_stream); pm = _lib.ptr(m)
for k in range(200): lib.mq_qnet_dropout_mask(pm, 64 * 512, 0.2, 1, k, st)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(3000): lib.mq_qnet_dropout_mask(pm, 64 * 512, 0.2, 1, k, st)
t = time.perf_counter() - t0; torch.cuda.synchronize()
print(f"mq_qnet_dropout_mask (tiny kernel through ctypes): host {t / 3000 * 1e6:.1f} us per launch")
pobs = _lib.ptr(obs)
for name, call in (("mq_env_reset", lambda: lib.mq_env_reset(env._h, None, None, pobs, None, st)),
                   ("mq_env_unpack_rmap", None)):
    if call is None:
        continue
    for k in range(50): call()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(1000): call()
    t = time.perf_counter() - t0; torch.cuda.synchronize()
    print(f"{name}: host {t / 1000 * 1e6:.1f} us per launch")
