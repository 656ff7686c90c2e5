"""Quick kernel-only timing of the log-mel kernel (resident inputs, CUDA events)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_crnn_b200 import feature

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kernel = sys.argv[2] if len(sys.argv) > 2 else "auto"          # auto | fp32 | tc
S = 180 * 44100
x = torch.empty(n_clips, 2, S, device="cuda").normal_(0, 0.1)
out = torch.empty(n_clips, feature.n_frames(S), 80, device="cuda")
for _ in range(3):
    feature.mbe_device(x, out=out, kernel=kernel)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
iters = 10
ev[0].record()
for _ in range(iters):
    feature.mbe_device(x, out=out, kernel=kernel)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / iters
frames = n_clips * 2 * feature.n_frames(S)
byts = x.numel() * 4 + out.numel() * 4
print(json.dumps(dict(kernel=kernel, n_clips=n_clips, ms=ms, frames_per_s=frames / ms * 1e3, GBps=byts / ms / 1e6,
                      audio_s_per_s=n_clips * 180 / ms * 1e3, frac_hbm=byts / ms / 1e6 / 6453.4)))
