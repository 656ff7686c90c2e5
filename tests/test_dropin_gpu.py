"""The nn.Module drop-ins behave like the reference classes: same state_dict keys/shapes, same outputs
for the same checkpoint, autograd + torch optimizers work, Lightning hooks run."""
import os

import numpy as np
import pytest
import torch

from oracle import crnn_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built_lib):
    assert torch.cuda.is_available()
    from sed_crnn_b200 import crnn_lightning, sed, modules
    return crnn_lightning, sed, modules


def test_state_dict_keys_match_reference_fixture(mods, golden_dir):
    cl, sed, _ = mods
    g = np.load(os.path.join(golden_dir, "crnn_fork_lightning.npz"))
    ref_keys = sorted(k[2:] for k in g.files if k.startswith("w."))
    m = cl.TimePooledCRNN(dropout=0.0)
    sd = m.state_dict()
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(g["w." + k].shape), k
    assert (m.T_out, m._flat) == (8, 640)
    g2 = np.load(os.path.join(golden_dir, "crnn_fork_sedpy.npz"))
    m2 = sed.TimePooledCRNN(conv_channels=32, dropout=0.0)
    assert sorted(m2.state_dict().keys()) == sorted(k[2:] for k in g2.files if k.startswith("w."))
    assert m2.flat == 32 * 40


def test_reference_checkpoint_one_step_with_torch_adam(mods, golden_dir):
    """load the reference's weights, run loss.backward() + clip_grad_norm_ + torch.optim.Adam exactly the
    way Lightning drives the module, compare with what the reference produced."""
    cl, _, _ = mods
    g = np.load(os.path.join(golden_dir, "crnn_fork_lightning.npz"))
    m = cl.TimePooledCRNN(dropout=0.0)
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
    m = m.cuda().train()
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    logits = m(x)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["logits0"], atol=2e-5)
    sd = m.state_dict()
    for k in ("conv_stack.1.running_mean", "conv_stack.5.running_var", "conv_stack.9.running_mean"):
        np.testing.assert_allclose(sd[k].cpu().numpy(), g["after_fwd." + k], rtol=1e-5, atol=1e-6)
    loss = cl.FocalBCELoss()(logits, y)
    assert abs(loss.item() - float(g["loss0"])) < 1e-6
    opt.zero_grad()
    loss.backward()
    for name, p in m.named_parameters():
        ref = g["grad." + name]
        if name in ("conv_stack.0.bias", "conv_stack.4.bias", "conv_stack.8.bias"):
            continue            # true gradient is 0 (BatchNorm follows); both sides hold rounding residue
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-6), name
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    assert abs(gn.item() - float(g["gnorm"])) < 1e-4 * float(g["gnorm"])
    opt.step()
    with torch.no_grad():
        p1 = torch.sigmoid(m(x)).cpu().numpy()
    assert np.abs(p1 - g["probs1"]).max() <= 1e-3
    assert np.array_equal(p1 > 0.5, g["probs1"] > 0.5)
    # the checkpoint was taken after the reference constructor's dry run (crnn_lightning.py:54-56 runs the
    # conv stack once in train mode): num_batches_tracked starts at 1
    assert int(m.state_dict()["conv_stack.1.num_batches_tracked"]) == 3


def test_sedpy_run_epoch_and_fused_optimizer(mods):
    _, sed, modules = mods
    torch.manual_seed(0)
    m = sed.TimePooledCRNN(conv_channels=32, dropout=0.5).cuda()
    g = torch.Generator().manual_seed(1)
    data = [(torch.randn(16, 1, 40, 64, generator=g), (torch.rand(16, 8, 1, generator=g) < 0.2).float()) for _ in range(3)]
    opt = modules.FusedClipAdam(m, lr=1e-3)
    l0, preds, labels = sed.run_epoch(m, data, sed.BCEWithLogitsLoss(), opt)
    l1, _, _ = sed.run_epoch(m, data, sed.BCEWithLogitsLoss(), opt)
    lv, pv, _ = sed.run_epoch(m, data, sed.BCEWithLogitsLoss())
    assert preds.shape == (48, 8, 1) and labels.shape == (48, 8, 1)
    assert np.isfinite([l0, l1, lv]).all() and l1 < l0


def test_lightning_hooks_and_device_metrics(mods):
    cl, _, _ = mods
    from oracle import metrics_ref as M
    torch.manual_seed(0)
    lm = cl.CRNNLightning(fold_id=1, art_dir="/tmp/unused").cuda()
    assert set(lm.hparams.keys()) >= {"fold_id", "lr", "weight_decay", "dropout"}
    cfgd = lm.configure_optimizers()
    opt = cfgd["optimizer"]
    g = torch.Generator().manual_seed(2)
    lm.train()
    for _ in range(3):
        batch = (torch.randn(32, 1, 40, 64, generator=g).cuda(), (torch.rand(32, 8, 1, generator=g) < 0.3).float().cuda())
        opt.zero_grad()
        loss = lm.training_step(batch, 0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(lm.parameters(), 1.0)
        opt.step()
    preds = torch.cat(lm._buf["train"]["preds"]).cpu().numpy()
    trues = torch.cat(lm._buf["train"]["trues"]).cpu().numpy()
    lm.on_train_epoch_end()
    O = (preds > 0.5).astype(np.uint8)
    assert lm.track["f1_1s_tr"][0] == M.f1_overall_1sec(O, trues.astype(np.uint8), 5)
    assert lm.track["er_1s_tr"][0] == M.er_overall_1sec(O, trues.astype(np.uint8), 5) or np.isnan(lm.track["er_1s_tr"][0])
    assert lm.track["f1_fr_tr"][0] == M.f1_overall_framewise(O, trues.astype(np.uint8))
    lm.eval()
    with torch.no_grad():
        lm.validation_step(batch, 0)
    lm.on_validation_epoch_end()
    assert "val_er_1s" in lm.logged and len(lm.track["loss_val"]) == 1


def test_fresh_module_has_the_reference_constructor_side_effects(mods):
    """crnn_lightning.py:54-56 / sed.py:94-98 push zeros through the conv stack in train mode inside
    __init__, so a fresh reference model has running_var = 0.9, running_mean = 0.1 * conv bias and
    num_batches_tracked = 1; the drop-ins reproduce that state."""
    cl, sed, _ = mods
    for m, bns, convs in ((cl.TimePooledCRNN(), [1, 5, 9], [0, 4, 8]),):
        sd = m.state_dict()
        for b, c in zip(bns, convs):
            assert int(sd[f"conv_stack.{b}.num_batches_tracked"]) == 1
            assert torch.allclose(sd[f"conv_stack.{b}.running_var"], torch.full((16,), 0.9))
            assert torch.allclose(sd[f"conv_stack.{b}.running_mean"], 0.1 * sd[f"conv_stack.{c}.bias"], atol=1e-7)
    m2 = sed.TimePooledCRNN(conv_channels=32)
    assert int(m2.state_dict()["bns.2.num_batches_tracked"]) == 1


def test_cpu_call_fails_loudly(mods):
    cl, _, _ = mods
    m = cl.TimePooledCRNN()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 1, 40, 64))


def test_device_prefetcher_orders_and_overlaps(mods):
    from sed_crnn_b200.parallel import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [(torch.randn(8, 1, 40, 64, generator=g).pin_memory(), torch.rand(8, 8, 1, generator=g).pin_memory())
            for _ in range(5)]
    seen = []
    pf = DevicePrefetcher(iter(host))
    for x, y, k in pf:
        seen.append((x.clone(), y.clone()))
        pf.release(k)
    assert len(seen) == 5
    for (xh, yh), (xd, yd) in zip(host, seen):
        assert torch.equal(xd.cpu(), xh) and torch.equal(yd.cpu(), yh)


def test_device_prefetcher_without_release_and_with_a_partial_last_batch(mods):
    """(a) a consumer that never calls release() must not have its buffer overwritten while work that reads
    it is still queued; (b) a shape change (last partial batch) allocates a new buffer that must be ordered after the
    current stream."""
    from sed_crnn_b200.parallel import DevicePrefetcher
    g = torch.Generator().manual_seed(1)
    host = [(torch.randn(b, 1, 40, 64, generator=g).pin_memory(), torch.rand(b, 8, 1, generator=g).pin_memory())
            for b in (64, 64, 64, 64, 64, 17)]
    sums = []
    for x, y, k in DevicePrefetcher(iter(host)):              # no release(): the fallback ordering must hold
        big = x.repeat(8, 1, 1, 1)
        for _ in range(20):                                    # keep the current stream busy reading x
            big = big * 1.0000001
        sums.append((x.double().sum() + 0 * big.sum().double(), y.double().sum()))
    assert len(sums) == 6
    for (xh, yh), (sx, sy) in zip(host, sums):
        assert abs(sx.item() - xh.double().sum().item()) < 1e-6 and abs(sy.item() - yh.double().sum().item()) < 1e-9


def test_fused_clip_adam_state_round_trips_through_a_checkpoint(mods):
    """optimizer.state_dict() must carry the Adam moments and the step (they live in the engine's flat
    buffers); a resumed run must continue exactly like an uninterrupted one, and a new optimizer starts from zero."""
    import io
    _, sed, modules = mods
    g = torch.Generator().manual_seed(3)
    data = [(torch.randn(8, 1, 40, 64, generator=g).cuda(), (torch.rand(8, 8, 1, generator=g) < 0.2).float().cuda())
            for _ in range(4)]
    loss_fn = sed.BCEWithLogitsLoss()

    def step(m, opt, xb, yb):
        opt.zero_grad()
        loss = loss_fn(m(xb), yb)
        loss.backward()
        opt.step()

    torch.manual_seed(0)
    a = sed.TimePooledCRNN(conv_channels=32, dropout=0.0).cuda()
    oa = modules.FusedClipAdam(a, lr=1e-3, max_norm=1.0)
    a.train()
    for xb, yb in data[:2]:
        step(a, oa, xb, yb)
    buf = io.BytesIO()
    torch.save({"model": a.state_dict(), "opt": oa.state_dict()}, buf)
    sd = oa.state_dict()
    assert sd["sedb200_flat"]["step"] == 2 and sd["sedb200_flat"]["exp_avg"].abs().max() > 0
    for xb, yb in data[2:]:
        step(a, oa, xb, yb)
    # resume in a fresh module + optimizer
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    torch.manual_seed(123)
    b = sed.TimePooledCRNN(conv_channels=32, dropout=0.0).cuda()
    b.load_state_dict(ck["model"])
    ob = modules.FusedClipAdam(b, lr=1e-3, max_norm=1.0)
    b.train()
    b(data[0][0])                                            # engine exists before the optimizer state arrives
    assert not b.engine.exp_avg.any() and b.engine.step_count == 0
    ob.load_state_dict(ck["opt"])
    assert b.engine.step_count == 2
    b.load_state_dict(ck["model"])                            # the probe forward above moved the BN running stats
    for xb, yb in data[2:]:
        step(b, ob, xb, yb)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka


def test_fused_clip_adam_fast_path_equals_general_route(mods):
    """FusedClipAdam skips the per-tensor gradient copies when every p.grad still aliases the flat buffer the backward
    pass returned; cloning the gradients first forces the general route -- both must produce the same parameters."""
    _, sed, modules = mods
    g = torch.Generator().manual_seed(5)
    data = [(torch.randn(8, 1, 40, 64, generator=g).cuda(), (torch.rand(8, 8, 1, generator=g) < 0.2).float().cuda())
            for _ in range(3)]
    loss_fn = sed.BCEWithLogitsLoss()
    outs = []
    for force_general in (False, True):
        torch.manual_seed(0)
        m = sed.TimePooledCRNN(conv_channels=32, dropout=0.0).cuda().train()
        opt = modules.FusedClipAdam(m, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
        took_fast = []
        for xb, yb in data:
            opt.zero_grad()
            loss_fn(m(xb), yb).backward()
            flat = m._last_flat_grad
            took_fast.append(all(p.grad is not None and flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + 4 * flat.numel()
                                 for p in m.parameters()))
            if force_general:
                for p in m.parameters():
                    p.grad = p.grad.clone()
            opt.step()
        outs.append((took_fast, {k: v.clone() for k, v in m.state_dict().items()}))
    assert all(outs[0][0]), "autograd did not adopt the returned gradient views: the fast path is never taken"
    for k in outs[0][1]:
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k
