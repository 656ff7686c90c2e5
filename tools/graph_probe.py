"""How much of the C2 step is launch gaps?  Captures one engine.train_step in a CUDA graph (with the seed / Adam step
frozen at capture time -- a TIMING probe only, the arithmetic of a replay is not a valid training step) and compares
replays with eager steps on the same device-resident batch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sed_crnn_b200 import config, engine

cfg = config.PRESETS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
B = 128
eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=1)
eng.init_default(0)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(cfg.input_shape(B), device="cuda", generator=g)
y = (torch.rand(cfg.target_shape(B), device="cuda", generator=g) < 0.2).float()
for _ in range(5):
    eng.train_step(x, y)
torch.cuda.synchronize()


def timed(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timed(lambda: eng.train_step(x, y))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    eng.train_step(x, y)
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    eng.train_step(x, y)
graph = timed(gr.replay)
print(json.dumps({"config": sys.argv[1] if len(sys.argv) > 1 else "c2", "eager_ms": eager, "graph_replay_ms": graph,
                  "gap_ms": eager - graph}))
