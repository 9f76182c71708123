"""Stand-alone tcgen05/TMEM/TMA bf16 GEMM (csrc/gemm_tc.cuh) through the C-ABI vs torch (fp32 matmul of the same
bf16-rounded operands).  fp32 accumulation in TMEM: agreement to ~1e-5 relative of the row scale; the tolerance below
also covers a different accumulation order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, bn, splits=1, seed=0):
    from dqn_marl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    A = torch.randn((M, K), generator=g, device="cuda:0").to(torch.bfloat16)
    B = torch.randn((N, K), generator=g, device="cuda:0").to(torch.bfloat16)
    C = torch.full((M, N), float("nan"), device="cuda:0")
    ws = torch.empty((max(splits, 1) * M * N,), device="cuda:0") if splits > 1 else None
    import ctypes as Ct
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = A.float() @ B.float().t()
    scale = ref.abs().max().item()
    if True:
        _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), None, M, N, K, bn, splits, _lib.ptr(ws), st), "mq_gemm_bf16")
        torch.cuda.synchronize()
        err = (C - ref).abs().max().item()
        assert err <= 2e-3 * scale + 1e-3, (M, N, K, bn, splits, err, scale)
    if splits == 1:          # bf16-only output: TMA-store epilogue when N is a multiple of the tile width
        Cb = torch.full((M + 3, N), 768.0, device="cuda:0").to(torch.bfloat16)
        _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(B), None, _lib.ptr(Cb), M, N, K, bn, 1, None, st), "mq_gemm_bf16")
        torch.cuda.synchronize()
        assert torch.all(Cb[M:].float() == 768.0), "rows past M were written"
        err = (Cb[:M].float() - ref).abs().max().item()
        assert err <= 1e-2 * scale + 1e-3, (M, N, K, bn, "bf16", err, scale)


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 128, 256, 128), (256, 512, 1024, 128), (1000, 384, 576, 128),
                                      (495, 64, 288, 64), (4096, 64, 1152, 64), (777, 32, 576, 32), (512, 15488, 512, 128),
                                      (128, 512, 15488, 128), (256, 512, 1024, 256), (300, 700, 576, 256), (512, 15488, 512, 256),
                                      (256, 256, 64, 512), (512, 512, 1024, 512), (300, 700, 576, 512), (640, 15488, 512, 512),
                                      (900, 384, 576, 384), (128, 128, 256, 384),
                                      (256, 256, 64, 1256), (512, 512, 1024, 1256), (300, 700, 576, 1256), (640, 15488, 512, 1256),
                                      (128, 128, 256, 1128), (1000, 384, 576, 1128), (4096, 512, 15488, 1256),
                                      (128, 32, 64, 2032), (777, 32, 64, 2032), (60000, 32, 128, 2032), (495616, 32, 64, 2032)])
def test_gemm_matches_torch(M, N, K, bn):
    _run(M, N, K, bn)


@pytest.mark.parametrize("M,N,K,bn,splits", [(576, 128, 8192, 128, 8), (288, 64, 4000, 64, 5), (512, 1024, 4096, 128, 3), (384, 512, 15488, 256, 2), (512, 512, 15488, 512, 4), (384, 640, 4096, 384, 3),
                                             (512, 512, 15488, 1256, 4), (384, 640, 4096, 1128, 3)])
def test_gemm_split_k(M, N, K, bn, splits):
    _run(M, N, K, bn, splits)


def _stream():
    import ctypes as Ct
    return Ct.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 64, 1), (128, 128, 256, 1), (512, 1024, 4096, 1), (512, 15488, 512, 1),
                                          (256, 384, 1000, 1), (576, 128, 8192, 8), (512, 256, 4100, 3)])
def test_gemm_tn_matches_torch(M, N, K, splits):
    """MN-major operands (K = row index): C = At^T Bt, the weight-gradient form."""
    from dqn_marl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda:0"); g.manual_seed(M + N + K)
    At = torch.randn((K, M), generator=g, device="cuda:0").to(torch.bfloat16)
    Bt = torch.randn((K, N), generator=g, device="cuda:0").to(torch.bfloat16)
    C = torch.full((M, N), float("nan"), device="cuda:0")
    ws = torch.empty((splits * M * N,), device="cuda:0") if splits > 1 else None
    _lib.check(lib.mq_gemm_bf16_tn(_lib.ptr(At), _lib.ptr(Bt), _lib.ptr(C), M, N, K, splits, _lib.ptr(ws), _stream()), "mq_gemm_bf16_tn")
    torch.cuda.synchronize()
    ref = At.float().t() @ Bt.float()
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (M, N, K, splits, err, scale)


def _conv_ref(X, Wk, Cin, Cout, flip):
    """X [B][11][11][Cin], Wk [Cout][9*Cin] with taps (kh, kw, c) -> [B*121][Cout] by torch conv2d on the same bf16 values."""
    w = Wk.float().reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)          # [Cout][Cin][kh][kw]
    if flip:
        w = w.flip(2, 3)
    y = torch.nn.functional.conv2d(X.float().permute(0, 3, 1, 2), w, padding=1)
    return y.permute(0, 2, 3, 1).reshape(-1, Cout)


@pytest.mark.parametrize("out", ["f32", "bf16"])
@pytest.mark.parametrize("B,Cin,Cout,bn,flip", [(1, 64, 128, 128, 0), (5, 64, 128, 128, 0), (64, 64, 128, 128, 1), (37, 128, 64, 64, 1),
                                                (16, 64, 32, 32, 1), (33, 32, 64, 64, 0), (300, 64, 128, 128, 0),
                                                (1, 64, 128, 0, 0), (700, 64, 128, 0, 0), (333, 128, 64, 0, 1), (450, 64, 32, 0, 1),
                                                (301, 32, 64, 0, 0)])
def test_implicit_conv_matches_torch(B, Cin, Cout, bn, flip, out):
    """3x3/pad-1 convolution as an implicit GEMM over shifted, zero-filled 4-D TMA boxes (no im2col buffer); bn = 0 is the
    persistent kernel, whose bf16 output goes through the shared-memory + TMA-store epilogue."""
    from dqn_marl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda:0"); g.manual_seed(B * 7 + Cin)
    X = torch.randn((B, 11, 11, Cin), generator=g, device="cuda:0").to(torch.bfloat16)
    Wk = (torch.randn((Cout, 9 * Cin), generator=g, device="cuda:0") / 8).to(torch.bfloat16)
    Y = torch.full((B * 121, Cout), float("nan"), device="cuda:0") if out == "f32" else None
    Yb = torch.full((B * 121 + 7, Cout), 768.0, device="cuda:0").to(torch.bfloat16) if out == "bf16" else None    # 7 guard rows
    _lib.check(lib.mq_conv3x3_bf16(_lib.ptr(X), _lib.ptr(Wk), _lib.ptr(Y), _lib.ptr(Yb), B, Cin, Cout, flip, bn, _stream()), "mq_conv3x3_bf16")
    torch.cuda.synchronize()
    ref = _conv_ref(X, Wk, Cin, Cout, flip)
    scale = ref.abs().max().item()
    if out == "f32":
        err = (Y - ref).abs().max().item()
        assert err <= 2e-3 * scale + 1e-3, (B, Cin, Cout, bn, flip, err, scale)
    else:
        assert torch.all(Yb[B * 121:].float() == 768.0), "rows past the last sample were written"
        err = (Yb[:B * 121].float() - ref).abs().max().item()
        assert err <= 1e-2 * scale + 1e-3, (B, Cin, Cout, bn, flip, err, scale)


@pytest.mark.parametrize("B,Cin,Cout,splits", [(1, 64, 128, 1), (9, 64, 128, 1), (100, 64, 128, 7), (64, 128, 64, 4), (257, 64, 128, 29),
                                               (3, 32, 64, 1), (130, 32, 64, 9)])
def test_conv_wgrad_matches_torch(B, Cin, Cout, splits):
    """dW[(kh,kw,c)][n] = sum over samples and pixels of X[b, i+kh-1, j+kw-1, c] * dY[b, i, j, n] (MN-major implicit GEMM)."""
    from dqn_marl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda:0"); g.manual_seed(B + Cout)
    X = torch.randn((B, 11, 11, Cin), generator=g, device="cuda:0").to(torch.bfloat16)
    dY = torch.randn((B * 121, Cout), generator=g, device="cuda:0").to(torch.bfloat16)
    dW = torch.full((9 * Cin, Cout), float("nan"), device="cuda:0")
    ws = torch.empty((splits * (9 * Cin + 1) * Cout,), device="cuda:0")
    has_spare_row = (9 * Cin) % 128 != 0
    db = torch.full((Cout,), float("nan"), device="cuda:0")
    if not has_spare_row:      # 9 Cin is a multiple of 128: no spare operand row, the bias gradient must be refused, not made up
        with pytest.raises(_lib.MqError, match="spare"):
            _lib.check(lib.mq_conv3x3_wgrad_bf16(_lib.ptr(X), _lib.ptr(dY), _lib.ptr(dW), _lib.ptr(db), B, Cin, Cout, splits, _lib.ptr(ws),
                                                 _stream()), "mq_conv3x3_wgrad_bf16")
        db = None
    _lib.check(lib.mq_conv3x3_wgrad_bf16(_lib.ptr(X), _lib.ptr(dY), _lib.ptr(dW), _lib.ptr(db), B, Cin, Cout, splits, _lib.ptr(ws), _stream()),
               "mq_conv3x3_wgrad_bf16")
    torch.cuda.synchronize()
    if has_spare_row:          # bias gradient = column sums of dY, from the spare operand row of ones
        ref_b = dY.float().sum(0)
        assert (db - ref_b).abs().max().item() <= 2e-3 * ref_b.abs().max().item() + 1e-3
    cols = torch.nn.functional.unfold(X.float().permute(0, 3, 1, 2), 3, padding=1)         # [B][Cin*9][121], (c, kh, kw) order
    cols = cols.reshape(B, Cin, 9, 121).permute(0, 3, 2, 1).reshape(B * 121, 9 * Cin)      # rows (b, pixel), cols (tap, c)
    ref = cols.t() @ dY.float()
    err = (dW - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (B, Cin, Cout, splits, err, scale)
