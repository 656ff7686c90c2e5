"""The recurrent scan kernels on their own (sedb200_gru_scan_fwd / _bwd through the C ABI) against torch.nn.GRU in
float64 -- every kernel family (H <= 16 warp-resident, H = 32 broadcast, H = 64 / 128 split, generic) at short,
BASELINE-C1 (256) and BASELINE-C5 (2,048) recurrence lengths.  Reference call sites: nn.GRU(bidirectional=True,
batch_first=True) in crnn_lightning.py:61-62,71 and sed.py:101,111.

nn.GRU is driven with an identity input projection (x = [gi_fwd | gi_rev], W_ih = [I 0] / [0 I], b_ih = 0), so its
input IS the pre-activation tensor the scan consumes and autograd's d(x) IS dgi."""
import pytest
import torch

import parity_util as PU

pytestmark = pytest.mark.gpu


def reference(gi, whh, bhh, dout):
    """gi [B,T,2,3H], whh [2,3H,H], bhh [2,3H], dout [B,T,2H] (all float64, CPU) -> out, dgi, dgh."""
    B, T, _, H3 = gi.shape
    H = H3 // 3
    gru = torch.nn.GRU(2 * H3, H, bidirectional=True, batch_first=True).double()
    eye = torch.eye(H3, dtype=torch.float64)
    zero = torch.zeros(H3, H3, dtype=torch.float64)
    with torch.no_grad():
        gru.weight_ih_l0.copy_(torch.cat([eye, zero], 1)); gru.weight_ih_l0_reverse.copy_(torch.cat([zero, eye], 1))
        gru.bias_ih_l0.zero_(); gru.bias_ih_l0_reverse.zero_()
        gru.weight_hh_l0.copy_(whh[0]); gru.weight_hh_l0_reverse.copy_(whh[1])
        gru.bias_hh_l0.copy_(bhh[0]); gru.bias_hh_l0_reverse.copy_(bhh[1])
    x = gi.reshape(B, T, 2 * H3).clone().requires_grad_(True)
    out, _ = gru(x)
    out.backward(dout)
    dgi = x.grad.reshape(B, T, 2, H3)
    # dgh = gradient w.r.t. (W_hh h_prev + b_hh): equal to dgi for the r and z gates, dgi_n * r for the n gate
    o = out.detach()
    hprev_f = torch.cat([torch.zeros(B, 1, H, dtype=torch.float64), o[:, :-1, :H]], 1)
    hprev_r = torch.cat([o[:, 1:, H:], torch.zeros(B, 1, H, dtype=torch.float64)], 1)
    dgh = dgi.clone()
    for d, hp in ((0, hprev_f), (1, hprev_r)):
        r = torch.sigmoid(gi[:, :, d, :H] + hp @ whh[d, :H].t() + bhh[d, :H])
        dgh[:, :, d, 2 * H:] = dgi[:, :, d, 2 * H:] * r
    return o, dgi, dgh, gru


CASES = [(H, T) for H in (8, 16, 32, 64, 128) for T in (8, 256, 2048)] + [(24, 40), (48, 64)]


@pytest.mark.parametrize("H,T", CASES)
def test_scan_forward_backward_vs_float64_gru(built_lib, H, T):
    from sed_crnn_b200 import _lib
    L = built_lib
    B = 3
    g = torch.Generator().manual_seed(1000 * H + T)
    gi = torch.randn(B, T, 2, 3 * H, generator=g)
    whh = (torch.rand(2, 3 * H, H, generator=g) * 2 - 1) / H ** 0.5             # nn.GRU default init range
    bhh = (torch.rand(2, 3 * H, generator=g) * 2 - 1) / H ** 0.5
    dout = torch.randn(B, T, 2 * H, generator=g)
    want_out, want_dgi, want_dgh, _ = reference(gi.double(), whh.double(), bhh.double(), dout.double())

    dev = "cuda"
    gi_d, whh_d, bhh_d, dout_d = gi.to(dev), whh.to(dev), bhh.to(dev), dout.to(dev)
    out = torch.full((B, T, 2 * H), float("nan"), device=dev)
    gates = torch.empty(B, T, 2, 4 * H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.sedb200_gru_scan_fwd(gi_d.data_ptr(), whh_d.data_ptr(), bhh_d.data_ptr(), out.data_ptr(),
                                      gates.data_ptr(), B, T, H, st))
    dgi = torch.full((B, T, 2, 3 * H), float("nan"), device=dev)
    dgh = torch.full((B, T, 2, 3 * H), float("nan"), device=dev)
    fused = bool(L.sedb200_gru_scan_fused_bias_grads(H))
    part_b = torch.zeros(B, 2, 2, 3 * H, device=dev)
    _lib.check(L.sedb200_gru_scan_bwd(dout_d.data_ptr(), out.data_ptr(), gates.data_ptr(), whh_d.data_ptr(),
                                      dgi.data_ptr(), dgh.data_ptr(), part_b.data_ptr() if fused else None, B, T, H, st))
    torch.cuda.synchronize()
    e_out = (out.cpu().double() - want_out).abs().max().item()
    sc_i, sc_h = want_dgi.abs().max().item(), want_dgh.abs().max().item()
    e_dgi = (dgi.cpu().double() - want_dgi).abs().max().item() / sc_i
    e_dgh = (dgh.cpu().double() - want_dgh).abs().max().item() / sc_h
    e_b = 0.0
    if fused:
        pb = part_b.cpu().double().sum(0)                                          # [ih|hh][dir][3H]
        wb = torch.stack([want_dgi.sum((0, 1)), want_dgh.sum((0, 1))])
        e_b = (pb - wb).abs().max().item() / wb.abs().max().item()
    PU.report(f"gru_scan_H{H}_T{T}", out_maxabs=e_out, dgi_rel=e_dgi, dgh_rel=e_dgh, bias_partials_rel=e_b)
    assert e_out <= 1e-5, e_out
    assert e_dgi <= 5e-5 and e_dgh <= 5e-5, (e_dgi, e_dgh)
    assert e_b <= 5e-5, e_b


def test_scan_argument_errors(built_lib):
    L = built_lib
    assert L.sedb200_gru_scan_fwd(None, None, None, None, None, 1, 1, 8, None) != 0
    assert b"null" in L.sedb200_last_error()
