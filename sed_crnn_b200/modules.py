"""nn.Module drop-ins for the reference's model / loss classes, backed by CRNNEngine.

The reference's boundary for this path is plain Python: `nn.Module`s called by Lightning's fit loop
(crnn_lightning.py:157-200) or by `sed.run_epoch` (sed.py:128-141).  These classes keep that surface --
class names, constructor signatures, attributes, `state_dict` keys and shapes, default initialisation --
and route `forward` / `backward` to the sm_100a kernels:

  * parameters live in real `nn.Conv2d / nn.BatchNorm2d / nn.GRU / nn.Linear` containers (never called),
    so `state_dict()` / `load_state_dict()` / checkpoints / default init are PyTorch's own, key for key;
  * their storage is re-pointed into the engine's flat parameter buffer; `.to()/.cuda()` is handled by
    re-flattening lazily on the next forward;
  * `forward` is one `torch.autograd.Function` over the whole network: `loss.backward()`,
    `clip_grad_norm_` and any `torch.optim` optimizer work unchanged (FusedClipAdam below is the fast
    flat-buffer alternative).

There is no CPU implementation: calling a module whose parameters are not on a CUDA device raises.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .config import CRNNConfig, FORK, SEDPY
from .engine import CRNNEngine, LOSS_KINDS


# ----------------------------------------------------------------------------------------------- autograd glue
class _CRNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module._engine
        ctx.module = module
        ctx.x = x
        ctx.training = module.training
        module._fwd_count += 1
        ctx.seed = eng.seed + module._fwd_count
        ctx.token = module._fwd_count
        logits = eng.forward(x, training=module.training, seed=ctx.seed)
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        module, eng = ctx.module, ctx.module._engine
        if not ctx.training:
            raise RuntimeError("backward through an eval-mode (running-statistics) forward is not implemented; "
                               "call .train() as the reference's training loops do")
        if ctx.token != module._fwd_count:
            raise RuntimeError("backward must follow the forward that produced the graph (activations of one "
                               "forward are kept at a time)")
        eng._last_seed = ctx.seed
        dx = torch.empty_like(ctx.x) if ctx.needs_input_grad[1] else None
        eng.backward(ctx.x, dlogits.contiguous(), dx=dx)
        flat = eng.grads.clone()                      # autograd owns what it is handed; the engine buffer is reused
        module._last_flat_grad = flat                 # FusedClipAdam recognises gradients that still alias this buffer
        grads = []
        for (name, shape, off, sub) in module._slots:
            n = int(np.prod(shape))
            v = flat[off:off + n].view(shape)
            grads.append(v[sub] if sub is not None else v)
        return (None, dx, *grads)


class EngineBackedCRNN(nn.Module):
    """Base: subclasses create the parameter containers and call `_bind(cfg, slots, bn_modules)` where
    slots = [(parameter, canonical_name, sub_index_or_None)]."""

    def _bind(self, cfg: CRNNConfig, slots, bn_modules, engine_kwargs=None, dry_run_stats=True):
        if dry_run_stats:
            # The reference constructors push a zero tensor through the conv stack in TRAIN mode to probe
            # shapes (crnn_lightning.py:54-56, sed.py:94-98).  That forward updates every BatchNorm: the
            # conv output is its bias everywhere (variance 0), BN maps it to beta = 0, so each block sees
            # zeros again.  Reproduce the resulting buffers instead of running a CPU forward.
            convs = [p for p, name, _ in slots if name.startswith("conv") and name.endswith(".bias")]
            with torch.no_grad():
                for bn, bias in zip(bn_modules, convs):
                    bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * bias.detach())
                    bn.running_var.mul_(1 - bn.momentum)
                    bn.num_batches_tracked += 1
        self.cfg = cfg
        self._engine = None
        self._engine_kwargs = dict(engine_kwargs or {})
        self._param_slots = slots
        self._bn_modules = bn_modules
        self._fwd_count = 0
        self._slots = None

    # ---- flat-buffer aliasing
    def _ensure_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = CRNNEngine(self.cfg, device=device, **self._engine_kwargs)
            spec = {name: (shape, off) for name, shape, off in self._engine.specs}
            self._slots = []
            for p, cname, sub in self._param_slots:
                shape, off = spec[cname]
                self._slots.append((cname, shape, off, sub))
        eng = self._engine
        views, bnv = eng.views(), eng.bn_views()
        with torch.no_grad():
            for (p, cname, sub) in self._param_slots:
                dst = views[cname] if sub is None else views[cname][sub]
                if p.data_ptr() != dst.data_ptr():
                    dst.copy_(p.data.to(device=device, dtype=torch.float32))
                    p.data = dst
            for i, bn in enumerate(self._bn_modules):
                for attr in ("running_mean", "running_var"):
                    dst = bnv[f"bn{i}.{attr}"]
                    cur = getattr(bn, attr)
                    if cur.data_ptr() != dst.data_ptr():
                        dst.copy_(cur.to(device=device, dtype=torch.float32))
                        setattr(bn, attr, dst)
        return eng

    def forward(self, x):
        p0 = self._param_slots[0][0]
        if not p0.is_cuda or not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__} runs on a CUDA (B200) device only -- move the module and its "
                               "input with .cuda(); there is no CPU fallback")
        eng = self._ensure_engine(p0.device)
        eng.desc.dropout = float(self._dropout_p()) if self.training else 0.0
        x = x.contiguous().float()
        params = [p for p, _, _ in self._param_slots]
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            out = _CRNNFunction.apply(self, x, *params)
        else:
            self._fwd_count += 1
            out = eng.forward(x, training=self.training, seed=eng.seed + self._fwd_count).clone()
        if self.training:
            for bn in self._bn_modules:
                bn.num_batches_tracked += 1
        return out

    def _dropout_p(self) -> float:
        return self.cfg.dropout

    @property
    def engine(self) -> CRNNEngine:
        return self._ensure_engine(self._param_slots[0][0].device)


def _gru_slots(gru: nn.GRU, layer: int, cidx: int):
    out = []
    for d, sfx in enumerate(("", "_reverse")):
        out += [(getattr(gru, f"weight_ih_l{layer}{sfx}"), f"gru{cidx}.w_ih", d),
                (getattr(gru, f"weight_hh_l{layer}{sfx}"), f"gru{cidx}.w_hh", d),
                (getattr(gru, f"bias_ih_l{layer}{sfx}"), f"gru{cidx}.b_ih", d),
                (getattr(gru, f"bias_hh_l{layer}{sfx}"), f"gru{cidx}.b_hh", d)]
    return out


# ----------------------------------------------------------------------------------------------- Lightning variant
class TimePooledCRNN(EngineBackedCRNN):
    """Drop-in for crnn_lightning.TimePooledCRNN (crnn_lightning.py:41-73): same ctor, same attributes
    (`conv_stack`, `gru1`, `gru2`, `d1`, `d2`, `T_out`, `_flat`), same state_dict keys, x [B,1,40,T] ->
    logits [B,T/8,1].  `cfg` generalises the hard-wired train_constants.py values."""

    def __init__(self, dropout=0.4, cfg: CRNNConfig | None = None, **engine_kwargs):
        super().__init__()
        cfg = FORK if cfg is None else cfg
        cfg = CRNNConfig(**{**cfg.__dict__, "dropout": float(dropout)})
        if len(cfg.gru_units) != 2 or len(cfg.dense_units) != 1:
            raise ValueError("TimePooledCRNN has exactly gru1, gru2, d1, d2 (use CRNN(cfg) for other depths)")
        self.conv_stack = nn.Sequential()
        in_c, slots, bns = cfg.in_ch, [], []
        for i, pool in enumerate(cfg.pool):
            conv, bn = nn.Conv2d(in_c, cfg.conv_ch, 3, padding=1), nn.BatchNorm2d(cfg.conv_ch)
            self.conv_stack.append(conv)
            self.conv_stack.append(bn)
            self.conv_stack.append(nn.ReLU())
            self.conv_stack.append(nn.MaxPool2d((1, pool)))
            slots += [(conv.weight, f"conv{i}.weight", None), (conv.bias, f"conv{i}.bias", None),
                      (bn.weight, f"bn{i}.weight", None), (bn.bias, f"bn{i}.bias", None)]
            bns.append(bn)
            in_c = cfg.conv_ch
        self.conv_stack.append(nn.Dropout(dropout))
        self.T_out, self._flat = cfg.seq_len_out, cfg.flat
        self.gru1 = nn.GRU(self._flat, cfg.gru_units[0], bidirectional=True, batch_first=True)
        self.gru2 = nn.GRU(2 * cfg.gru_units[0], cfg.gru_units[1], bidirectional=True, batch_first=True)
        self.d1 = nn.Linear(2 * cfg.gru_units[1], cfg.dense_units[0])
        self.d2 = nn.Linear(cfg.dense_units[0], cfg.n_classes)
        slots += _gru_slots(self.gru1, 0, 0) + _gru_slots(self.gru2, 0, 1)
        slots += [(self.d1.weight, "dense0.weight", None), (self.d1.bias, "dense0.bias", None),
                  (self.d2.weight, "dense1.weight", None), (self.d2.bias, "dense1.bias", None)]
        self._bind(cfg, slots, bns, engine_kwargs)

    def _dropout_p(self):
        return self.conv_stack[-1].p


# ----------------------------------------------------------------------------------------------- sed.py variant
class SedTimePooledCRNN(EngineBackedCRNN):
    """Drop-in for sed.TimePooledCRNN (sed.py:82-112): `convs`, `bns`, `pools`, `drop`, `gru`
    (nn.GRU(num_layers=2, hidden 32)), `fc`, attribute `flat`."""

    def __init__(self, conv_channels=128, dropout=0.5, cfg: CRNNConfig | None = None, **engine_kwargs):
        super().__init__()
        base = SEDPY if cfg is None else cfg
        cfg = CRNNConfig(**{**base.__dict__, "conv_ch": int(conv_channels), "dropout": float(dropout)})
        if len(set(cfg.gru_units)) != 1 or cfg.dense_units:
            raise ValueError("sed.TimePooledCRNN has one nn.GRU(num_layers=n) and a single fc layer")
        self.convs, self.bns, self.pools = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        ch, slots = cfg.in_ch, []
        for i, p in enumerate(cfg.pool):
            self.convs.append(nn.Conv2d(ch, cfg.conv_ch, 3, padding=1))
            self.bns.append(nn.BatchNorm2d(cfg.conv_ch))
            self.pools.append(nn.MaxPool2d(kernel_size=(1, p)))
            ch = cfg.conv_ch
        self.drop = nn.Dropout(dropout)
        self.flat = cfg.flat
        hid = cfg.gru_units[0]
        self.gru = nn.GRU(self.flat, hid, num_layers=len(cfg.gru_units), batch_first=True, bidirectional=True)
        self.fc = nn.Linear(2 * hid, cfg.n_classes)
        for i in range(len(cfg.pool)):
            slots += [(self.convs[i].weight, f"conv{i}.weight", None), (self.convs[i].bias, f"conv{i}.bias", None),
                      (self.bns[i].weight, f"bn{i}.weight", None), (self.bns[i].bias, f"bn{i}.bias", None)]
        for layer in range(len(cfg.gru_units)):
            slots += _gru_slots(self.gru, layer, layer)
        slots += [(self.fc.weight, "dense0.weight", None), (self.fc.bias, "dense0.bias", None)]
        self._bind(cfg, slots, list(self.bns), engine_kwargs)

    def _dropout_p(self):
        return self.drop.p


# ----------------------------------------------------------------------------------------------- general factory
class CRNN(EngineBackedCRNN):
    """Any CRNNConfig (e.g. the BASELINE.json SEDnet configs) with canonical parameter names."""

    def __init__(self, cfg: CRNNConfig, **engine_kwargs):
        super().__init__()
        self.convs, self.bns, self.grus, self.denses = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        ch, slots = cfg.in_ch, []
        for i, p in enumerate(cfg.pool):
            self.convs.append(nn.Conv2d(ch, cfg.conv_ch, 3, padding=1))
            self.bns.append(nn.BatchNorm2d(cfg.conv_ch))
            slots += [(self.convs[i].weight, f"conv{i}.weight", None), (self.convs[i].bias, f"conv{i}.bias", None),
                      (self.bns[i].weight, f"bn{i}.weight", None), (self.bns[i].bias, f"bn{i}.bias", None)]
            ch = cfg.conv_ch
        d = cfg.flat
        for i, h in enumerate(cfg.gru_units):
            self.grus.append(nn.GRU(d, h, bidirectional=True, batch_first=True))
            slots += _gru_slots(self.grus[i], 0, i)
            d = 2 * h
        for i, u in enumerate(list(cfg.dense_units) + [cfg.n_classes]):
            self.denses.append(nn.Linear(d, u))
            slots += [(self.denses[i].weight, f"dense{i}.weight", None), (self.denses[i].bias, f"dense{i}.bias", None)]
            d = u
        self.T_out, self.flat = cfg.seq_len_out, cfg.flat
        self._bind(cfg, slots, list(self.bns), engine_kwargs, dry_run_stats=False)


def get_model(name_or_cfg="fork", **kwargs):
    """The `get_model()` factory the upstream README names (README.md:44): returns the CRNN for a preset
    ("fork", "sedpy", "c1", "c2", "c5") or a CRNNConfig."""
    from .config import PRESETS
    if isinstance(name_or_cfg, CRNNConfig):
        return CRNN(name_or_cfg, **kwargs)
    if name_or_cfg == "fork":
        return TimePooledCRNN(**kwargs)
    if name_or_cfg == "sedpy":
        return SedTimePooledCRNN(**kwargs)
    return CRNN(PRESETS[name_or_cfg], **kwargs)


# ----------------------------------------------------------------------------------------------- losses
class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, kind, alpha, gamma, reduction):
        if not logits.is_cuda:
            raise RuntimeError("loss runs on a CUDA (B200) device only; there is no CPU fallback")
        lg = logits.contiguous().float()
        tg = targets.contiguous().float().expand_as(lg).contiguous()
        n = lg.numel()
        scale = float(n) if reduction == "sum" else 1.0
        out = torch.empty(1, device=lg.device)
        dlog = torch.empty_like(lg)
        scratch = torch.empty(1024, device=lg.device)
        with torch.cuda.device(lg.device):
            _lib.check(_lib.lib().sedb200_loss_fwd_bwd(kind, alpha, gamma, lg.data_ptr(), tg.data_ptr(), n, scale,
                                                       out.data_ptr(), None, dlog.data_ptr(), scratch.data_ptr(),
                                                       scratch.numel() * 4, _lib.current_stream_ptr()))
        ctx.save_for_backward(dlog)
        return out[0] * scale

    @staticmethod
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        return dlog * g, None, None, None, None, None


class FocalBCELoss(nn.Module):
    """Drop-in for crnn_lightning.FocalBCELoss (crnn_lightning.py:27-35)."""

    def __init__(self, alpha=.25, gamma=2., reduction="mean"):
        super().__init__()
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, logits, targets):
        return _LossFunction.apply(logits, targets, LOSS_KINDS["focal"], float(self.alpha), float(self.gamma),
                                   "mean" if self.reduction == "mean" else "sum")


class BCEWithLogitsLoss(nn.Module):
    """Drop-in for the nn.BCEWithLogitsLoss() of sed.py:160 (mean reduction)."""

    def forward(self, logits, targets):
        return _LossFunction.apply(logits, targets, LOSS_KINDS["bce"], 0.0, 0.0, "mean")


# ----------------------------------------------------------------------------------------------- optimizer
class FusedClipAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, weight_decay) [+ clip_grad_norm_(max_norm)] as ONE pass over the engine's flat
    buffers (crnn_lightning.py:195-197 + train_lightning.py:50; sed.py:159 with max_norm=None)."""

    def __init__(self, module: EngineBackedCRNN, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 max_norm=None):
        self.module = module
        super().__init__(list(module.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_norm = max_norm
        if module._engine is not None:             # a NEW optimizer starts from zero moments, like torch.optim.Adam
            module._engine.reset_optimizer()

    # The moments live in the engine's flat buffers, not in Optimizer.state; checkpoints must still carry them
    # (torch.optim.Adam's state round-trips through state_dict, and Lightning saves optimizer.state_dict()).
    def state_dict(self):
        eng = self.module.engine
        sd = super().state_dict()
        sd["sedb200_flat"] = {"exp_avg": eng.exp_avg.detach().cpu().clone(),
                              "exp_avg_sq": eng.exp_avg_sq.detach().cpu().clone(), "step": int(eng.step_count),
                              "layout": [(n, tuple(sh), int(off)) for n, sh, off in eng.specs]}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("sedb200_flat", None)
        super().load_state_dict(state_dict)
        if flat is None:
            raise KeyError("this optimizer state has no 'sedb200_flat' entry: it was not saved by FusedClipAdam")
        eng = self.module.engine
        if [(n, tuple(sh), int(off)) for n, sh, off in eng.specs] != [(n, tuple(sh), int(off)) for n, sh, off in flat["layout"]]:
            raise ValueError("optimizer state was saved for a different parameter layout")
        with torch.no_grad():
            eng.exp_avg.copy_(flat["exp_avg"].to(eng.device))
            eng.exp_avg_sq.copy_(flat["exp_avg_sq"].to(eng.device))
        eng.step_count = int(flat["step"])

    @torch.no_grad()
    def step(self, closure=None):
        eng = self.module.engine
        g = self.param_groups[0]
        eng.lr, eng.betas, eng.eps, eng.weight_decay = g["lr"], g["betas"], g["eps"], g["weight_decay"]
        eng.clip = float(self.max_norm) if self.max_norm else 0.0
        # Fast path: after zero_grad(set_to_none=True) + one backward, autograd adopted the views of the flat gradient
        # buffer the backward pass returned -- every p.grad still sits at its offset of that buffer, so the buffer IS
        # the flat gradient and no per-tensor copies are needed.  Anything else (accumulated gradients, hooks that
        # replaced a gradient, a missing one) takes the general route.
        flat = getattr(self.module, "_last_flat_grad", None)
        aliased = flat is not None and flat.device == eng.grads.device
        if aliased:
            base = flat.data_ptr()
            for (p, cname, sub), (_, shape, off, _) in zip(self.module._param_slots, self.module._slots):
                g = p.grad
                if g is None or not g.is_contiguous():
                    aliased = False
                    break
                n = 1
                for d in shape:
                    n *= d
                want = base + 4 * (off + (sub * (n // shape[0]) if sub is not None else 0))
                if g.data_ptr() != want:
                    aliased = False
                    break
        if aliased:
            keep, eng.grads = eng.grads, flat
            try:
                eng.optimizer_step(1)
            finally:
                eng.grads = keep
            return None
        views = eng.views(eng.grads)
        eng.grads.zero_()
        for (p, cname, sub) in self.module._param_slots:
            if p.grad is not None:
                (views[cname] if sub is None else views[cname][sub]).copy_(p.grad)
        eng.optimizer_step(1)
        return None
