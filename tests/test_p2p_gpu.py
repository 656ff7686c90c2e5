"""Fused peer-memory all-reduce + clip + Adam kernel (sedb200_p2p_allreduce_clip_adam).

Single GPU (world = 1: the exchange region is local, same kernel, same flag protocol) against the plain
sedb200_clip_adam path; the two-rank NVLink run lives in tests/mgpu_check.py (needs `gpurun --gpus 2`)."""
from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engines(**kw):
    from sed_crnn_b200 import config, engine
    cfg = replace(config.PRESETS["c2"], seq_len=32, dropout=0.0)
    a = engine.CRNNEngine(cfg, loss="bce", **kw)
    b = engine.CRNNEngine(cfg, loss="bce", grad_exchange="p2p", **kw)
    a.init_default(3)
    b.init_default(3)
    return cfg, a, b


@pytest.mark.parametrize("clip", [1.0, 0.0])
def test_p2p_world1_matches_clip_adam(built_lib, clip):
    cfg, a, b = _engines(clip=clip, weight_decay=1e-4)
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(4):
        x = torch.randn(cfg.input_shape(8), device="cuda", generator=g)
        y = (torch.rand(cfg.target_shape(8), device="cuda", generator=g) < 0.2).float()
        la, _ = a.train_step(x, y)
        lb, _ = b.train_step(x, y)
        assert torch.equal(la, lb)
        # the summed gradient IS the local gradient at world 1; the norm is folded in a different (fixed) order,
        # so the clip coefficient may differ in its last bit
        assert torch.equal(a.grads, b.grads)
        assert torch.allclose(a._scalars[1], b._scalars[1], rtol=1e-6, atol=0)
        assert torch.allclose(a.params, b.params, rtol=0, atol=2e-7), (a.params - b.params).abs().max()
        assert torch.allclose(a.exp_avg_sq, b.exp_avg_sq, rtol=1e-5, atol=1e-12)
    assert b.xch.status() == 0 and b.xch.seq == 4
    # the two halves of the exchange region alternate
    assert b.xch.next_grad_buffer().data_ptr() != b.grads.data_ptr()
    b.xch.close()


def test_p2p_rejects_bad_arguments(built_lib):
    import ctypes as C
    from sed_crnn_b200 import _lib
    L = _lib.lib()
    assert L.sedb200_p2p_region_bytes(0) == 0
    assert L.sedb200_p2p_region_bytes(1024) >= 1024 + 2 * 4096
    assert L.sedb200_p2p_grad_offset_bytes(1024, 1) - L.sedb200_p2p_grad_offset_bytes(1024, 0) == 4096
    tab = (C.c_void_p * 1)(None)
    rc = L.sedb200_p2p_allreduce_clip_adam(tab, 1, 0, 1024, 1, 1, None, None, None, None, 1e-3, .9, .999, 1e-8, 0., 1.,
                                           1., None, None, 0, None)
    assert rc == _lib.EINVAL
    rc = L.sedb200_p2p_allreduce_clip_adam(tab, 17, 0, 1024, 1, 1, None, None, None, None, 1e-3, .9, .999, 1e-8, 0.,
                                           1., 1., None, None, 0, None)
    assert rc == _lib.EINVAL


def test_p2p_silent_peer_aborts_the_step_and_raises(built_lib, monkeypatch):
    """on a flag-wait timeout the kernel used to continue and apply a stale peer buffer.  Now it raises the
    sticky status word, leaves parameters and Adam state untouched, writes gnorm = NaN, and the next host call raises.
    A silent peer is simulated on one GPU: a second exchange region nobody publishes into is listed as rank 1."""
    import ctypes as C
    from sed_crnn_b200 import _lib, parallel
    monkeypatch.setenv("SEDB200_P2P_TIMEOUT_MS", "50")
    L = _lib.lib()
    n = 4096
    xch = parallel.P2PGradExchange(n, torch.device("cuda"))
    nbytes = int(L.sedb200_p2p_region_bytes(n))
    fake, handle = C.c_void_p(), (C.c_ubyte * 64)()
    _lib.check(L.sedb200_p2p_region_alloc(nbytes, C.byref(fake), handle))
    xch.world, xch.table = 2, (C.c_void_p * 2)(xch.own, fake.value)
    params = torch.randn(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p0 = params.clone()
    xch.next_grad_buffer().fill_(0.5)
    gn = torch.zeros(1, device="cuda")
    xch.allreduce_clip_adam(params, m, v, step=1, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip=1.0,
                            gnorm_out=gn)
    torch.cuda.synchronize()
    assert torch.equal(params, p0) and not m.any() and not v.any()
    assert torch.isnan(gn).all()
    assert xch.status() == 1
    with pytest.raises(RuntimeError, match="did not publish"):
        xch.raise_if_failed(wait=True)
    with pytest.raises(RuntimeError, match="did not publish"):
        xch.allreduce_clip_adam(params, m, v, step=2, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                                clip=1.0, gnorm_out=gn)
    xch.world = 1
    xch.close()
    L.sedb200_p2p_region_free(fake)


def test_two_gpu_exchange_against_virtual_replica_oracle(built_lib):
    """tests/mgpu_check.py under torchrun when the box has at least two GPUs (skipped on a single-GPU box): ranks
    bit-identical, fused NVLink kernel == NCCL path, both == the 2-virtual-replica CPU oracle after one step."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
                        os.path.join(root, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{") and '"verdict"' in l][-1]
    assert json.loads(line)["verdict"] == "PASS"
