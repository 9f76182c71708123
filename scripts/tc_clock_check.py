"""SM clock and power while a tcgen05 kernel runs back to back for a few seconds (is the GEMM efficiency a clock effect?)."""
import ctypes as Ct, os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200 import _lib
lib = _lib.load()
B = 4096
st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
X = torch.randn(B, 11, 11, 64, device="cuda").bfloat16(); Wk = torch.randn(128, 576, device="cuda").bfloat16()
Y = torch.empty(B * 121, 128, device="cuda", dtype=torch.bfloat16)
A = torch.randn(B, 15488, device="cuda").bfloat16(); W1 = torch.randn(512, 15488, device="cuda").bfloat16()
C = torch.empty(B, 512, device="cuda"); ws = torch.empty(2 * B * 512, device="cuda")
def conv(): _lib.check(lib.mq_conv3x3_bf16(_lib.ptr(X), _lib.ptr(Wk), None, _lib.ptr(Y), B, 64, 128, 0, 0, st), "conv")
def fc1(): _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(W1), _lib.ptr(C), None, B, 512, 15488, 256, 2, _lib.ptr(ws), st), "gemm")
def mm(): torch.matmul(A, W1.t())
for name, fn, flop in (("conv3 fwd persistent", conv, 2.0 * B * 121 * 128 * 576), ("fc1 fwd 128x256", fc1, 2.0 * B * 512 * 15488), ("torch.matmul (cuBLAS) fc1 shape", mm, 2.0 * B * 512 * 15488)):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "100", "-i", "0"],
                         stdout=subprocess.PIPE, text=True)
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time(); e0.record()
    while time.time() - t0 < 2.5:
        for _ in range(200): fn()
        n += 200
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    p.terminate()
    out = p.stdout.read().strip().splitlines()
    mid = out[len(out) // 2:]
    print(f"{name}: {ms * 1e3:.1f} us, {flop / ms / 1e9:.0f} TFLOP/s; samples (late half): {mid[:6]}")
