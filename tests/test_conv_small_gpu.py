"""Small-channel direct 3x3 convolutions (conv_small.cu) against torch.nn.functional on the CPU (float64).

Covers the reference's shipped geometry (train_constants.py: 1 -> 16 -> 16 channels, 40 x 64/32/16), NCHW and
channels-last inputs, odd row counts (partial last band) and the row-sliced weight-gradient path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (K = in channels, N = out channels, B, H, W, nchw input)
CASES = [(1, 16, 3, 40, 64, True), (16, 16, 2, 40, 32, False), (16, 16, 2, 13, 16, False), (2, 64, 3, 24, 40, True),
         (32, 32, 2, 9, 16, False), (64, 64, 1, 24, 8, False), (2, 16, 2, 7, 8, True)]


def _run(K, N, B, H, W, nchw, L, _lib):
    g = torch.Generator().manual_seed(K * 1000 + N * 10 + W)
    x = torch.randn(B, K, H, W, generator=g)
    w = torch.randn(N, K, 3, 3, generator=g) * 0.2
    bias = torch.randn(N, generator=g)
    dy = torch.randn(B, N, H, W, generator=g)
    if nchw:
        xin = x.contiguous()
        strides = (K * H * W, W, 1, H * W)
    else:
        xin = x.permute(0, 2, 3, 1).contiguous()
        strides = (H * W * K, W * K, K, 1)
    dev = lambda t: t.float().cuda().contiguous()
    xd, wd, bd = dev(xin), dev(w), dev(bias)
    dyd = dev(dy.permute(0, 2, 3, 1))
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    if L.sedb200_conv3x3_small_supported(K, N, W, 0):
        y = torch.empty(B, H, W, N, device="cuda")
        _lib.check(L.sedb200_conv3x3_small(xd.data_ptr(), *strides, K, B, H, W, wd.data_ptr(), bd.data_ptr(), N, 0,
                                           y.data_ptr(), st))
        want = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
        out["fwd"] = (y.cpu().double() - want).abs().max().item() / want.abs().max().item()
    if L.sedb200_conv3x3_small_supported(N, K, W, 0) and K % 8 == 0:
        dx = torch.empty(B, H, W, K, device="cuda")
        _lib.check(L.sedb200_conv3x3_small(dyd.data_ptr(), H * W * N, W * N, N, 1, N, B, H, W, wd.data_ptr(), None, K, 1,
                                           dx.data_ptr(), st))
        want = F.conv_transpose2d(dy.double(), w.double(), padding=1).permute(0, 2, 3, 1)
        out["dgrad"] = (dx.cpu().double() - want).abs().max().item() / want.abs().max().item()
    if L.sedb200_conv3x3_small_supported(K, N, W, 1):
        nb = int(L.sedb200_conv3x3_small_wgrad_scratch_bytes(K, N, B, H))
        scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
        dw = torch.empty(N, K, 3, 3, device="cuda")
        _lib.check(L.sedb200_conv3x3_small_wgrad(dyd.data_ptr(), xd.data_ptr(), *strides, K, N, B, H, W, dw.data_ptr(),
                                                 scratch.data_ptr(), nb, st))
        xr = x.double().requires_grad_(False)
        wr = w.double().requires_grad_(True)
        F.conv2d(xr, wr, None, padding=1).backward(dy.double())
        out["wgrad"] = (dw.cpu().double() - wr.grad).abs().max().item() / wr.grad.abs().max().item()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("K,N,B,H,W,nchw", CASES)
def test_conv_small_matches_torch(built_lib, K, N, B, H, W, nchw):
    from sed_crnn_b200 import _lib
    out = _run(K, N, B, H, W, nchw, built_lib, _lib)
    assert out, "no kernel supports this case"
    for name, err in out.items():
        assert err <= 2e-5, (name, err)


def test_conv_small_rejects_unsupported(built_lib):
    from sed_crnn_b200 import _lib
    assert built_lib.sedb200_conv3x3_small_supported(16, 12, 64, 0) == 0          # N % 8
    assert built_lib.sedb200_conv3x3_small_supported(16, 16, 30, 0) == 0          # W % 4
    assert built_lib.sedb200_conv3x3_small_supported(128, 128, 8, 1) == 0         # too many (n, k) pairs
    x = torch.zeros(16, device="cuda")
    rc = built_lib.sedb200_conv3x3_small(x.data_ptr(), 1, 1, 1, 1, 16, 1, 1, 30, x.data_ptr(), None, 16, 0, x.data_ptr(), None)
    assert rc == _lib.ESHAPE
