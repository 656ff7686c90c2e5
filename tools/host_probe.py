"""Host-side enqueue cost of each C-ABI call of a training step (diagnostic, 1 GPU)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_crnn_b200 import config, engine, _lib

cfg = config.PRESETS[sys.argv[1] if len(sys.argv) > 1 else "c2"]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0)
eng.init_default(0)
x = torch.randn(cfg.input_shape(B), device="cuda"); y = (torch.rand(cfg.target_shape(B), device="cuda") < 0.2).float()
for _ in range(3): eng.train_step(x, y)
torch.cuda.synchronize()
L = eng.L
sync_each = len(sys.argv) > 3 and sys.argv[3] == 'sync'
acc = {}
orig = {}
for name in ("sedb200_crnn_forward", "sedb200_loss_fwd_bwd", "sedb200_crnn_head_fwd_bwd", "sedb200_crnn_backward", "sedb200_clip_adam"):
    f = getattr(L, name)
    def wrap(*a, _f=f, _n=name):
        t0 = time.perf_counter(); r = _f(*a); acc[_n] = acc.get(_n, 0.0) + time.perf_counter() - t0; return r
    orig[name] = f
    setattr(L, name, wrap)
N = 200
l0 = L.sedb200_launch_count()
t0 = time.perf_counter()
for _ in range(N):
    eng.train_step(x, y)
    if sync_each: torch.cuda.synchronize()
t_host = (time.perf_counter() - t0) / N * 1e3
launches = (L.sedb200_launch_count() - l0) / N
torch.cuda.synchronize()
print("host ms/step", round(t_host, 3), "launches/step", launches)
for k, v in acc.items(): print(" ", k, round(v / N * 1e3, 3), "ms")
