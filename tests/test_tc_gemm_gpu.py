"""tcgen05 GEMM building block (3-term bf16 split) vs float64 matmul, all three operand-major combinations."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(L, a, a_mn, b, b_mn, M, N, K, bias, split_k):
    from sed_crnn_b200 import _lib
    out = torch.full((M, N), float("nan"), device="cuda")
    nbytes = L.sedb200_gemm_tc_scratch_bytes(M, N, K)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.sedb200_gemm_tc(a.data_ptr(), a_mn, b.data_ptr(), b_mn, M, N, K,
                                 bias.data_ptr() if bias is not None else None, out.data_ptr(), int(split_k),
                                 scratch.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().double()


@pytest.mark.parametrize("M,N,K", [(32768, 192, 256), (1000, 64, 64), (128, 128, 64), (4096, 192, 64), (40000, 256, 192)])
def test_tn_projection(built_lib, M, N, K):
    g = torch.Generator().manual_seed(M)
    a, b, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    want = a.double() @ b.double().t() + bias.double()
    got = run(built_lib, a.cuda(), 0, b.cuda(), 0, M, N, K, bias.cuda(), False)
    assert (got - want).abs().max().item() / want.abs().max().item() < 2e-5


@pytest.mark.parametrize("M,N,K", [(32768, 256, 192), (512, 64, 192), (3000, 32, 24)])
def test_k_major_times_mn_major(built_lib, M, N, K):
    """dX = dgi @ W: A [M][K] K-major, B given as W [K][N] (MN-major)."""
    g = torch.Generator().manual_seed(N)
    a, w = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g) / K ** 0.5
    want = a.double() @ w.double()
    got = run(built_lib, a.cuda(), 0, w.cuda(), 1, M, N, K, None, False)
    assert (got - want).abs().max().item() / want.abs().max().item() < 2e-5


@pytest.mark.parametrize("M,N,K", [(192, 256, 32768), (192, 64, 32768), (96, 32, 4096), (16, 64, 1000 * 8), (256, 512, 2048)])
def test_at_b_split_k(built_lib, M, N, K):
    """dW = dgi^T @ X: both operands stored [K][rows] (MN-major), reduction over K rows with split-K."""
    g = torch.Generator().manual_seed(K)
    a, b = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    want = a.double().t() @ b.double()
    got = run(built_lib, a.cuda(), 1, b.cuda(), 1, M, N, K, None, True)
    assert (got - want).abs().max().item() / want.abs().max().item() < 2e-5
