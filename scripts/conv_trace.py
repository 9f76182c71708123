"""Throw-away: per-role barrier wait cycles of the persistent conv kernel (needs a -DMQ_CONV_TRACE build, MARL_B200_SO)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from dqn_marl_b200 import _lib
lib = _lib.load()
lib.mq_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
B = 4096
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, cin, cout, flip in [("conv3 fwd", 64, 128, 0), ("conv3 dgrad", 128, 64, 1), ("conv2 fwd", 32, 64, 0), ("conv2 dgrad", 64, 32, 1)]:
    X = torch.randn(B, 11, 11, cin, device="cuda").bfloat16()
    Wk = torch.randn(cout, 9 * cin, device="cuda").bfloat16()
    Y = torch.empty(B * 121, cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        _lib.check(lib.mq_conv3x3_bf16(_lib.ptr(X), _lib.ptr(Wk), None, _lib.ptr(Y), B, cin, cout, flip, 0, st), "conv")
    torch.cuda.synchronize()
    lib.mq_debug_conv_trace(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mq_conv3x3_bf16(_lib.ptr(X), _lib.ptr(Wk), None, _lib.ptr(Y), B, cin, cout, flip, 0, st), "conv")
    e1.record(); torch.cuda.synchronize()
    out = np.zeros(256 * 8, dtype=np.int64)
    lib.mq_debug_conv_trace(ctypes.c_void_p(out.ctypes.data), 0)
    t = out.reshape(256, 8)[:148]
    n = t[:, 5].mean()
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.1f} us; samples/CTA {n:.1f}; per-CTA mean cycles: producer wait empty {t[:,0].mean():.0f}, "
          f"mma wait tmem_empty {t[:,1].mean():.0f}, mma wait full {t[:,2].mean():.0f}, epi wait tmem_full {t[:,3].mean():.0f}, epi work {t[:,4].mean():.0f}")
