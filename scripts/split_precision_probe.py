"""Can error-compensated bf16 splits on tcgen05 reach the 1e-5 parity bar of the fp32 path?

An fp32 matrix X is written as hi + mid + lo (three bf16 planes, 24 mantissa bits together); the product A B^T is the sum
of the six leading plane products, which the EXISTING bf16 GEMM computes in one launch when the planes are concatenated
along K (A' = [hi|hi|mid|hi|lo|mid], B' = [hi|mid|hi|lo|hi|mid]): the fp32 accumulator in TMEM adds them.  What the probe
measures is the error of that accumulation (bf16 x bf16 products are exact in fp32) against a float64 product, next to
torch's fp32 matmul (the FFMA bar), for K = 15488 (fc1), 576 (conv3) and split-K 1 / 8 (shorter chains).
"""
import ctypes as Ct
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200 import _lib

lib = _lib.load()
torch.backends.cuda.matmul.allow_tf32 = False
st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def split3(x):
    hi = x.to(torch.bfloat16)
    r = x - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def gemm(A, B, splits, bn=128):
    m, k = A.shape
    n = B.shape[0]
    C = torch.empty((m, n), device="cuda", dtype=torch.float32)
    ws = torch.empty((splits * m * n,), device="cuda") if splits > 1 else None
    _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), None, m, n, k, bn, splits, _lib.ptr(ws), st), "gemm")
    torch.cuda.synchronize()
    return C


def report(name, C, ref):
    err = (C.double() - ref).abs()
    scale = ref.abs().mean()
    print(f"  {name:34s} max|err|/mean|ref| {float(err.max() / scale):.3e}   mean|err|/mean|ref| {float(err.mean() / scale):.3e}   "
          f"mean signed {float((C.double() - ref).mean() / scale):+.3e}")


torch.manual_seed(0)
for label, M, N, K, relu_a in [("fc1 forward shape, relu'd activations", 512, 512, 15488, True),
                               ("fc1 forward shape, signed", 512, 512, 15488, False),
                               ("conv3 shape, relu'd", 1024, 128, 576, True)]:
    A = torch.randn((M, K), device="cuda")
    if relu_a:
        A = A.clamp_min(0)
    B = (torch.rand((N, K), device="cuda") * 2 - 1) / K ** 0.5
    if relu_a:
        B = B + 0.3 / K ** 0.5          # a same-signed sum: the worst case for a truncating accumulator
    ref = A.double() @ B.double().t()
    print(label, f"M={M} N={N} K={K}")
    report("torch fp32 matmul", A @ B.t(), ref)
    ah, am, al = split3(A)
    bh, bm, bl = split3(B)
    report("bf16 hi only", gemm(ah, bh, 1), ref)
    A3 = torch.cat([ah, ah, am], 1).contiguous(); B3 = torch.cat([bh, bm, bh], 1).contiguous()
    report("3 products (hi,mid)", gemm(A3, B3, 1), ref)
    A6 = torch.cat([ah, ah, am, ah, al, am], 1).contiguous(); B6 = torch.cat([bh, bm, bh, bl, bh, bm], 1).contiguous()
    for sp in (1, 8, 32):
        report(f"6 products, split-K {sp}", gemm(A6, B6, sp), ref)
    # small terms first: the big hi*hi product lands on an accumulator that already holds the corrections
    A6r = torch.cat([am, al, ah, am, ah, ah], 1).contiguous(); B6r = torch.cat([bm, bh, bl, bh, bm, bh], 1).contiguous()
    report("6 products, small terms first", gemm(A6r, B6r, 1), ref)
    # interleaved per 64-wide K block (what an in-kernel split loop would do)
    kb = K // 64
    def inter(parts):
        return torch.stack([q.view(q.shape[0], kb, 64) for q in parts], 2).reshape(parts[0].shape[0], -1).contiguous()
    A6i = inter([ah, ah, am, ah, al, am]); B6i = inter([bh, bm, bh, bl, bh, bm])
    for sp in (1, 8):
        report(f"6 products interleaved, split-K {sp}", gemm(A6i, B6i, sp), ref)
