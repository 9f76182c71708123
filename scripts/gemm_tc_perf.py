"""Throughput of the stand-alone tcgen05 GEMM on the Q-network's shapes (B = 4096): TFLOP/s vs MEASURED_PEAKS bf16."""
import ctypes as Ct
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200 import _lib

lib = _lib.load()
B = 4096
M = B * 121
shapes = [("fc1 fwd", B, 512, 15488, 1128, 1), ("fc1 fwd", B, 512, 15488, 1256, 1), ("fc1 fwd bf16", B, 512, 15488, 1128, 1), ("fc1 fwd", B, 512, 15488, 1256, 2), ("fc1 dgrad bf16", B, 15488, 512, 128, 1), ("fc1 dgrad bf16", B, 15488, 512, 256, 1), ("fc1 dgrad bf16", B, 15488, 512, 384, 1), ("fc1 dgrad bf16", B, 15488, 512, 512, 1), ("fc1 dgrad bf16", B, 15488, 512, 1256, 1), ("fc1 dgrad bf16", B, 15488, 512, 1128, 1), ("fc2 fwd bf16", B, 256, 512, 128, 1), ("fc2 dgrad bf16", B, 512, 256, 128, 1), ("conv1 fwd", M, 32, 64, 32, 1), ("conv1 fwd", M, 32, 64, 2032, 1), ("fc1 fwd", B, 512, 15488, 1256, 2), ("fc1 fwd", B, 512, 15488, 1256, 4), ("fc1 fwd", B, 512, 15488, 1128, 2), ("fc1 fwd", B, 512, 15488, 1128, 4), ("fc1 dgrad", B, 15488, 512, 1256, 1), ("fc1 dgrad", B, 15488, 512, 1128, 1), ("fc1 fwd", B, 512, 15488, 128, 2), ("fc1 fwd", B, 512, 15488, 256, 2), ("fc1 fwd", B, 512, 15488, 256, 4), ("fc1 fwd", B, 512, 15488, 512, 4), ("fc1 fwd", B, 512, 15488, 384, 2), ("fc1 fwd", B, 512, 15488, 384, 3), ("fc1 dgrad", B, 15488, 512, 256, 1), ("fc1 dgrad", B, 15488, 512, 512, 1), ("fc1 dgrad", B, 15488, 512, 384, 1), ("fc1 fwd", B, 512, 15488, 128, 1), ("conv3 fwd", M, 128, 576, 128, 1), ("conv2 fwd", M, 64, 288, 64, 1),
          ("fc1 dgrad", B, 15488, 512, 128, 1), ("fc1 wgrad", 512, 15488, B, 128, 1), ("conv3 dgrad", M, 64, 1152, 64, 1),
          ("conv3 wgrad", 576, 128, M, 128, 96), ("conv2 wgrad", 288, 64, M, 64, 148), ("conv2 dgrad", M, 32, 576, 32, 1)]
peak = 1629.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["bf16_tflops"]
st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, m, n, k, bn, splits in shapes:
    A = torch.randn((m, k), device="cuda").to(torch.bfloat16)
    Bm = torch.randn((n, k), device="cuda").to(torch.bfloat16)
    bf16_out = name.endswith("bf16")          # bf16-only output (TMA-store epilogue), as the learner uses these shapes
    C = torch.empty((m, n), device="cuda", dtype=torch.bfloat16 if bf16_out else torch.float32)
    cf, cb = (None, _lib.ptr(C)) if bf16_out else (_lib.ptr(C), None)
    ws = torch.empty((splits * m * n,), device="cuda") if splits > 1 else None
    for _ in range(3):
        _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(Bm), cf, cb, m, n, k, bn, splits, _lib.ptr(ws), st), "gemm")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 10
    for _ in range(reps):
        lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(Bm), cf, cb, m, n, k, bn, splits, _lib.ptr(ws), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    print(f"{name:12s} M={m:7d} N={n:6d} K={k:7d} bn={bn:3d} splits={splits:3d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  {100*tf/peak:5.1f}% of {peak:.0f}")

print("--- implicit convolutions (one-sample-per-CTA kernel vs persistent kernel) and MN-major weight gradients")
def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, cin, cout, flip, bn in [("conv3 fwd", 64, 128, 0, 128), ("conv3 dgrad", 128, 64, 1, 64), ("conv2 fwd", 32, 64, 0, 64),
                                  ("conv2 dgrad", 64, 32, 1, 32)]:
    X = torch.randn((B, 11, 11, cin), device="cuda").to(torch.bfloat16)
    Wk = torch.randn((cout, 9 * cin), device="cuda").to(torch.bfloat16)
    Y = torch.empty((M, cout), device="cuda").to(torch.bfloat16)          # bf16 activations out, as in the Q-network
    fl = 2.0 * M * cout * 9 * cin
    for label, b in (("tile", bn), ("persistent", 0)):
        ms = timeit(lambda: _lib.check(lib.mq_conv3x3_bf16(_lib.ptr(X), _lib.ptr(Wk), None, _lib.ptr(Y), B, cin, cout, flip, b, st), "conv"))
        tf = fl / (ms * 1e-3) / 1e12
        print(f"{name:12s} {label:10s} Cin={cin:3d} Cout={cout:3d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  {100*tf/peak:5.1f}% of {peak:.0f}")
for name, cin, cout, splits in [("conv3 wgrad", 64, 128, 30), ("conv2 wgrad", 32, 64, 50)]:
    X = torch.randn((B, 11, 11, cin), device="cuda").to(torch.bfloat16)
    dY = torch.randn((M, cout), device="cuda").to(torch.bfloat16)
    dW = torch.empty((9 * cin, cout), device="cuda")
    ws = torch.empty((splits * 9 * cin * cout,), device="cuda")
    ms = timeit(lambda: _lib.check(lib.mq_conv3x3_wgrad_bf16(_lib.ptr(X), _lib.ptr(dY), _lib.ptr(dW), None, B, cin, cout, splits, _lib.ptr(ws), st), "wgrad"))
    tf = 2.0 * M * cout * 9 * cin / (ms * 1e-3) / 1e12
    print(f"{name:12s} implicit TN splits={splits:3d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  {100*tf/peak:5.1f}% of {peak:.0f}")
At = torch.randn((B, 512), device="cuda").to(torch.bfloat16)
Bt = torch.randn((B, 15488), device="cuda").to(torch.bfloat16)
C = torch.empty((512, 15488), device="cuda")
ms = timeit(lambda: _lib.check(lib.mq_gemm_bf16_tn(_lib.ptr(At), _lib.ptr(Bt), _lib.ptr(C), 512, 15488, B, 1, None, st), "tn"))
tf = 2.0 * 512 * 15488 * B / (ms * 1e-3) / 1e12
print(f"fc1 wgrad    MN-major TN            {ms:8.3f} ms  {tf:7.1f} TFLOP/s  {100*tf/peak:5.1f}% of {peak:.0f}")
