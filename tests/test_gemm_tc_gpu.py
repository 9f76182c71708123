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
    _lib.check(lib.mq_gemm_bf16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, bn, splits, _lib.ptr(ws), st), "mq_gemm_bf16")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (M, N, K, bn, splits, err, scale)


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 128, 256, 128), (256, 512, 1024, 128), (1000, 384, 576, 128),
                                      (495, 64, 288, 64), (4096, 64, 1152, 64), (777, 32, 576, 32), (512, 15488, 512, 128),
                                      (128, 512, 15488, 128)])
def test_gemm_matches_torch(M, N, K, bn):
    _run(M, N, K, bn)


@pytest.mark.parametrize("M,N,K,bn,splits", [(576, 128, 8192, 128, 8), (288, 64, 4000, 64, 5), (512, 1024, 4096, 128, 3)])
def test_gemm_split_k(M, N, K, bn, splits):
    _run(M, N, K, bn, splits)
