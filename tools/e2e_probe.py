"""Where does the end-to-end loop lose time against the device-resident loop?  (diagnostic, 1 GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_crnn_b200 import config, engine
from sed_crnn_b200.parallel import DevicePrefetcher

cfg = config.PRESETS["c2"]; B = 128
eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0)
eng.init_default(0)
g = torch.Generator().manual_seed(1)
xs_h = [torch.randn(cfg.input_shape(B), generator=g).pin_memory() for _ in range(4)]
ys_h = [(torch.rand(cfg.target_shape(B), generator=g) < 0.2).float().pin_memory() for _ in range(4)]
xs = [t.cuda() for t in xs_h]; ys = [t.cuda() for t in ys_h]
loss_h = torch.empty(1).pin_memory()
N = 30
def timed(fn):
    for _ in range(2): fn(4)
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(N); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e3
def resident(n):
    for i in range(n): eng.train_step(xs[i % 4], ys[i % 4])
def resident_sync(n):
    for i in range(n):
        l, _ = eng.train_step(xs[i % 4], ys[i % 4]); loss_h.copy_(l.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
def hb(n):
    for i in range(n): yield xs_h[i % 4], ys_h[i % 4]
def e2e(n):
    pf = DevicePrefetcher(hb(n))
    for xd, yd, k in pf:
        l, _ = eng.train_step(xd, yd); pf.release(k); loss_h.copy_(l.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
def e2e_nosync(n):
    pf = DevicePrefetcher(hb(n))
    for xd, yd, k in pf:
        l, _ = eng.train_step(xd, yd); pf.release(k); loss_h.copy_(l.reshape(1), non_blocking=True)
def host_only(n):
    # python + launch cost of a step with the GPU kept far behind? (enqueue time)
    t0 = time.perf_counter()
    for i in range(n): eng.train_step(xs[i % 4], ys[i % 4])
    host_only.t = (time.perf_counter() - t0) / n * 1e3
for name, fn in [("resident", resident), ("resident+sync", resident_sync), ("e2e", e2e), ("e2e_nosync", e2e_nosync)]:
    print(name, round(timed(fn), 4), "ms/step")
torch.cuda.synchronize(); host_only(30); print("host enqueue per step", round(host_only.t, 4), "ms"); torch.cuda.synchronize()
