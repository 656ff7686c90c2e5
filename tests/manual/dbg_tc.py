import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dataclasses import replace
from oracle import crnn_ref as R
from sed_crnn_b200 import config, engine
for preset, seq, B in (("c1", 32, 6), ("c2", 32, 8), ("c1", 256, 16)):
    outs = {}
    for tc in (False, True):
        rcfg = {**R.PRESETS[preset], "seq_len": seq}
        torch.manual_seed(0)
        ref = R.RefCRNN(**rcfg)
        eng = engine.CRNNEngine(replace(config.PRESETS[preset], seq_len=seq, dropout=0.0, tensor_cores=tc), loss="bce")
        eng.load_named({k: v.detach() for k, v in ref.canonical_named_params()})
        x, y = R.synth_batch(rcfg, B, seed=3)
        logits = eng.forward(x.cuda(), training=True).clone()
        loss, probs, dlog = eng.loss_and_grad(logits, y.cuda())
        eng.backward(x.cuda(), dlog)
        outs[tc] = (logits, {k: v.clone() for k, v in eng.views(eng.grads).items()})
    print(preset, seq, B, "logits", (outs[True][0] - outs[False][0]).abs().max().item())
    for k in outs[True][1]:
        a, b = outs[True][1][k], outs[False][1][k]
        print("   %-14s rel %.3e   max|g| %.3e" % (k, ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), b.abs().max().item()))
