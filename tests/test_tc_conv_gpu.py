"""tcgen05 convolution building block (through the C ABI) vs a float64 torch.nn.functional.conv2d oracle.
Checks the 3-term bf16 split really delivers fp32-grade results (error ~1e-6 of the output scale,
where single-pass bf16 would sit at ~4e-3)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built_lib):
    assert torch.cuda.is_available()
    return built_lib


@pytest.fixture(params=["0", "1"], ids=["boxes-per-tap", "halo-boxes"])
def halo(request, monkeypatch):
    """conv_tc_kernel's operand staging: one box per tap (default) or one halo box per tap column (SEDB200_CONV_HALO=1,
    read by the library at every launch)."""
    monkeypatch.setenv("SEDB200_CONV_HALO", request.param)
    return request.param


def run_tc(L, x_nhwc, w, bias, dgrad):
    from sed_crnn_b200 import _lib
    B, H, W, Ck = x_nhwc.shape
    Cout, Cin = w.shape[:2]
    Cn = Cin if dgrad else Cout
    out = torch.empty(B, H, W, Cn, device="cuda")
    nbytes = L.sedb200_conv3x3_tc_scratch_bytes(B, H, W, Cin, Cout)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.sedb200_conv3x3_tc(x_nhwc.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                    out.data_ptr(), B, H, W, Cin, Cout, int(dgrad), scratch.data_ptr(), nbytes,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 8, 64, 128), (1, 40, 32, 128, 128), (3, 16, 4, 128, 128),
                                            (2, 7, 16, 64, 256), (5, 256, 8, 128, 128), (1, 3, 128, 64, 128),
                                            (21, 256, 8, 64, 128),           # 336 tiles: > 2 per CTA
                                            (3, 24, 8, 64, 128), (2, 40, 16, 128, 128),   # halo boxes: ragged last tile
                                            (2, 9, 64, 64, 128)])            # W = 64: the halo rows do not fit
def test_forward_matches_float64_conv(L, halo, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    got = run_tc(L, x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), b.cuda(), False).cpu().double()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 8, 128, 64), (1, 40, 32, 128, 128), (2, 16, 4, 256, 128),
                                            (2, 40, 16, 128, 128), (3, 24, 8, 128, 128)])
def test_dgrad_matches_autograd(L, halo, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(7 + H)
    x = torch.randn(B, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5))
    dy = torch.randn(B, Cout, H, W, generator=g)
    F.conv2d(x, w.double(), None, padding=1).backward(dy.double())
    want = x.grad.permute(0, 2, 3, 1)
    got = run_tc(L, dy.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), None, True).cpu().double()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 8, 128, 128), (3, 10, 16, 128, 128), (2, 16, 4, 256, 128),
                                            (1, 5, 32, 128, 256), (7, 64, 8, 128, 128)])
def test_wgrad_matches_autograd(L, B, H, W, Cin, Cout):
    from sed_crnn_b200 import _lib
    g = torch.Generator().manual_seed(11 + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(B, Cout, H, W, generator=g)
    F.conv2d(x.double(), w, None, padding=1).backward(dy.double())
    want = w.grad
    dw = torch.empty(Cout, Cin, 3, 3, device="cuda")
    nbytes = L.sedb200_conv3x3_wgrad_tc_scratch_bytes(B, H, W, Cin, Cout)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    _lib.check(L.sedb200_conv3x3_wgrad_tc(dyd.data_ptr(), xd.data_ptr(), dw.data_ptr(), B, H, W, Cin, Cout,
                                          scratch.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (dw.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err


# ---------------------------------------------------------------------------------------------------------------------
# The PLANE-NATIVE launches the CRNN flow actually runs (fp16 planes: forward = fp16 pass + e4m3 correction pass on one
# tile, gradients = one fp16 pass on tile pairs, halo boxes in all three kernels), through the test hook
# sedb200_conv3x3_planes_test.  Two kinds of input: small dyadic values, for which every product and every fp32 partial
# sum is exact -- so the result must EQUAL the float64 convolution and a wrong tap / row / channel shows up as an O(1)
# error, not as something a gradient gate could absorb -- and random values with the tolerance of the operand format.
def run_planes(L, mode, a, b, w, B, H, W, Cin, Cout):
    from sed_crnn_b200 import _lib
    shape = {0: (B, H, W, Cout), 1: (B, H, W, Cin), 2: (Cout, Cin, 3, 3)}[mode]
    out = torch.empty(*shape, device="cuda")
    nbytes = L.sedb200_conv3x3_planes_test_scratch_bytes(B, H, W, Cin, Cout)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.sedb200_conv3x3_planes_test(a.data_ptr(), b.data_ptr() if b is not None else None,
                                             w.data_ptr() if w is not None else None, out.data_ptr(), B, H, W, Cin, Cout,
                                             mode, scratch.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().double()


def dyadic(shape, g, levels, denom):
    return torch.randint(-levels, levels + 1, shape, generator=g).float() / denom


@pytest.fixture(params=[("1", "0"), ("1", "2"), ("0", "0")], ids=["halo", "halo-pairs", "boxes-per-tap"])
def staging(request, monkeypatch):
    halo, tpi = request.param
    monkeypatch.setenv("SEDB200_CONV_HALO", halo)
    monkeypatch.setenv("SEDB200_WGRAD_HALO", halo)
    monkeypatch.setenv("SEDB200_CONV_TPI", tpi)
    return request.param


PLANE_SHAPES = [(2, 40, 16, 128, 128), (3, 24, 8, 128, 128), (2, 40, 4, 128, 128), (1, 40, 32, 128, 128),
                (2, 9, 64, 128, 128), (3, 33, 8, 128, 256), (2, 16, 4, 256, 128)]


@pytest.mark.parametrize("B,H,W,Cin,Cout", PLANE_SHAPES)
@pytest.mark.parametrize("exact", [True, False], ids=["dyadic", "random"])
def test_plane_native_forward(L, staging, exact, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(31 * H + W + B)
    if exact:
        x, w = dyadic((B, Cin, H, W), g, 3, 4.0), dyadic((Cout, Cin, 3, 3), g, 4, 16.0)
    else:
        x, w = torch.randn(B, Cin, H, W, generator=g), torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    want = F.conv2d(x.double(), w.double(), None, padding=1).permute(0, 2, 3, 1)
    got = run_planes(L, 0, x.permute(0, 2, 3, 1).contiguous().cuda(), None, w.cuda(), B, H, W, Cin, Cout)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= (1e-6 if exact else 1e-4), err          # fp16 x fp16 + e4m3 corrections: ~2^-15 per product


@pytest.mark.parametrize("B,H,W,Cin,Cout", PLANE_SHAPES)
@pytest.mark.parametrize("exact", [True, False], ids=["dyadic", "random"])
def test_plane_native_data_gradient(L, staging, exact, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(17 * H + W + B)
    if exact:
        dy, w = dyadic((B, Cout, H, W), g, 3, 4.0), dyadic((Cout, Cin, 3, 3), g, 4, 16.0)
    else:
        dy, w = torch.randn(B, Cout, H, W, generator=g), torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), None, padding=1).backward(dy.double())
    want = x.grad.permute(0, 2, 3, 1)
    got = run_planes(L, 1, dy.permute(0, 2, 3, 1).contiguous().cuda(), None, w.cuda(), B, H, W, Cin, Cout)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= (1e-6 if exact else 2e-3), err          # ONE fp16 pass: 2^-11 per operand


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 8, 128, 128), (3, 10, 16, 128, 128), (2, 16, 4, 256, 128),
                                            (1, 5, 32, 128, 256), (7, 64, 8, 128, 128), (2, 33, 4, 128, 128)])
@pytest.mark.parametrize("exact", [True, False], ids=["dyadic", "random"])
def test_plane_native_weight_gradient(L, staging, exact, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(13 * H + W + B)
    if exact:
        x, dy = dyadic((B, Cin, H, W), g, 3, 4.0), dyadic((B, Cout, H, W), g, 3, 4.0)
    else:
        x, dy = torch.randn(B, Cin, H, W, generator=g), torch.randn(B, Cout, H, W, generator=g)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, padding=1).backward(dy.double())
    want = w.grad
    got = run_planes(L, 2, dy.permute(0, 2, 3, 1).contiguous().cuda(), x.permute(0, 2, 3, 1).contiguous().cuda(), None,
                     B, H, W, Cin, Cout)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= (1e-6 if exact else 2e-3), err
