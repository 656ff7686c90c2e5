"""CRNNEngine -- the B200 training / inference engine behind the drop-in modules.

Owns the flat parameter / gradient / Adam-state buffers and the activation workspace (all PyTorch
tensors: PyTorch is the allocator and the stream provider, nothing more) and drives libsedb200.so:

    forward -> loss(+dlogits) -> backward -> [NCCL all-reduce of the flat gradient] -> clip + Adam

which is the arithmetic of crnn_lightning.py:157-163 + train_lightning.py:50 + crnn_lightning.py:195-197
(Lightning variant) or sed.py:134-137,159 (plain-torch variant).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import CRNNConfig

LOSS_KINDS = {"bce": 0, "focal": 1}


class CRNNEngine:
    def __init__(self, cfg: CRNNConfig, device="cuda", *, loss="focal", alpha=0.25, gamma=2.0,
                 lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, clip=1.0, seed=0,
                 process_group=None, grad_exchange: str = "nccl", cuda_graph: bool = False):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CRNNEngine runs on a CUDA device only (no CPU fallback)")
        self.L = _lib.lib()
        self.desc = cfg.desc()
        _lib.check(self.L.sedb200_crnn_validate(C.byref(self.desc)))
        self.loss_kind, self.alpha, self.gamma = LOSS_KINDS[loss], float(alpha), float(gamma)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.clip = float(clip) if clip else 0.0
        self.seed, self.step_count = int(seed), 0
        self.pg = process_group
        self.specs = cfg.tensor_specs()
        n = cfg.n_param_floats()
        kw = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(n, **kw)
        self.grads = torch.zeros(n, **kw)
        self.exp_avg = torch.zeros(n, **kw)
        self.exp_avg_sq = torch.zeros(n, **kw)
        nb = int(self.L.sedb200_crnn_bn_state_floats(C.byref(self.desc)))
        self.bn_state = torch.zeros(nb, **kw)
        c = cfg.conv_ch
        for i in range(len(cfg.pool)):
            self.bn_state[2 * i * c + c: 2 * (i + 1) * c] = 1.0          # running_var = 1
        self.num_batches_tracked = 0
        self._scratch = torch.zeros(4096, **kw)
        self._scalars = torch.zeros(4, **kw)                              # loss, gnorm
        self._counts = torch.zeros(13, dtype=torch.int64, device=self.device)
        self._ws = None
        self._ws_batch = 0
        self._bufs = {}
        self._last_seed = self.seed
        # cuda_graph: single-process train_step calls are captured once per (input buffers, hyper-parameters) and
        # replayed; the per-step dropout seed and Adam bias corrections live in a device-side step state
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self._graph_misses = 0
        self._step_state = torch.zeros(int(self.L.sedb200_step_state_bytes()) // 8, dtype=torch.int64, device=self.device)
        self._state_step = 0                        # the step counter the device-side state holds
        self.graph_replays = 0
        self.launches_per_graph_step = 0
        # "nccl": dist.all_reduce + clip_adam kernels;  "p2p": one fused kernel over NVLink peer memory
        if grad_exchange not in ("nccl", "p2p"):
            raise ValueError("grad_exchange must be 'nccl' or 'p2p'")
        # two-layer dense heads run as one fused forward + loss + backward kernel inside train_step
        self.fused_head = bool(self.L.sedb200_crnn_head_supported(C.byref(self.desc)))
        self.xch = None
        if grad_exchange == "p2p":
            from . import parallel
            self.xch = parallel.P2PGradExchange(n, self.device, process_group)
            self.grads = self.xch.next_grad_buffer()

    # ------------------------------------------------------------------ parameters
    def views(self, flat: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
        flat = self.params if flat is None else flat
        out = {}
        for name, shape, off in self.specs:
            n = 1
            for s in shape:
                n *= s
            out[name] = flat[off:off + n].view(shape)
        return out

    def bn_views(self) -> dict[str, torch.Tensor]:
        c, out = self.cfg.conv_ch, {}
        for i in range(len(self.cfg.pool)):
            out[f"bn{i}.running_mean"] = self.bn_state[2 * i * c: 2 * i * c + c]
            out[f"bn{i}.running_var"] = self.bn_state[2 * i * c + c: 2 * (i + 1) * c]
        return out

    @torch.no_grad()
    def load_named(self, tensors: dict) -> None:
        """Copy tensors given under canonical names.  GRU tensors may be given per direction as
        `gru{i}.f.w_ih` / `gru{i}.r.w_ih` (the oracle's naming) or stacked as `gru{i}.w_ih`."""
        v = self.views()
        v.update(self.bn_views())
        for name, t in tensors.items():
            t = torch.as_tensor(t)
            parts = name.split(".")
            if len(parts) == 3 and parts[1] in ("f", "r"):
                dst = v[f"{parts[0]}.{parts[2]}"][0 if parts[1] == "f" else 1]
            else:
                dst = v[name]
            dst.copy_(t.to(self.device, torch.float32).reshape(dst.shape))

    @torch.no_grad()
    def init_default(self, seed: int = 0) -> None:
        """PyTorch-default-style random init (uniform +-1/sqrt(fan_in); BatchNorm weight 1 / bias 0),
        drawn on the host from `seed` so that every rank gets identical weights."""
        g = torch.Generator().manual_seed(seed)
        host, fan = {}, 1
        for name, shape, _ in self.specs:
            layer, leaf = name.split(".")
            if layer.startswith("bn"):
                host[name] = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
                continue
            if layer.startswith("conv") and leaf == "weight":
                fan = shape[1] * 9
            elif layer.startswith("gru"):
                fan = shape[1] // 3                                   # nn.GRU: 1/sqrt(hidden_size) for all
            elif layer.startswith("dense") and leaf == "weight":
                fan = shape[1]
            host[name] = (torch.rand(shape, generator=g) * 2 - 1) / (fan ** 0.5)
        self.load_named(host)

    def reset_optimizer(self) -> None:
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0

    # ------------------------------------------------------------------ plumbing
    def _workspace(self, batch: int):
        if self._ws is None or self._ws_batch != batch:
            nbytes = int(self.L.sedb200_crnn_workspace_bytes(C.byref(self.desc), batch))
            if nbytes == 0:
                _lib.check(self.L.sedb200_crnn_validate(C.byref(self.desc)))
                raise _lib.Sedb200Error(_lib.ESHAPE, f"batch {batch}: " + self.L.sedb200_last_error().decode(errors="replace"))
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_batch = batch
        return self._ws

    def _buf(self, key, shape):
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._bufs[key] = b
        return b

    def _check_x(self, x):
        if not (x.is_cuda and x.dtype == torch.float32):
            raise TypeError("x must be a CUDA float32 tensor")
        want = self.cfg.input_shape(x.shape[0])
        if tuple(x.shape) != want:
            raise ValueError(f"x has shape {tuple(x.shape)}, config expects {want}")
        return x.contiguous()

    # ------------------------------------------------------------------ compute
    def forward(self, x: torch.Tensor, training: bool = False, logits: torch.Tensor | None = None,
                seed: int | None = None, skip_head: bool = False) -> torch.Tensor:
        """`skip_head`: stop after the GRU stack (the fused head kernel takes over); returns None."""
        x = self._check_x(x)
        self._last_seed = self.seed + self.step_count if seed is None else int(seed)
        B = x.shape[0]
        ws = self._workspace(B)
        if logits is None and not skip_head:
            logits = self._buf(("logits", B), self.cfg.target_shape(B))
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_crnn_forward(
                C.byref(self.desc), self.params.data_ptr(), self.bn_state.data_ptr(), x.data_ptr(), B,
                int(training), self._last_seed, ws.data_ptr(), ws.numel(),
                None if skip_head else logits.data_ptr(), _lib.current_stream_ptr()))
        if training:
            self.num_batches_tracked += 1
        return logits

    def loss_and_grad(self, logits, targets, grad_scale: float = 1.0, want_grad: bool = True):
        """-> (loss [device scalar view], probs, dlogits)"""
        n = logits.numel()
        targets = targets.contiguous()
        if targets.shape != logits.shape or targets.dtype != torch.float32 or not targets.is_cuda:
            raise ValueError("targets must be CUDA float32 with the logits' shape")
        probs = self._buf(("probs", tuple(logits.shape)), logits.shape)
        dlog = self._buf(("dlogits", tuple(logits.shape)), logits.shape) if want_grad else None
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_loss_fwd_bwd(
                self.loss_kind, self.alpha, self.gamma, logits.data_ptr(), targets.data_ptr(), n, float(grad_scale),
                self._scalars.data_ptr(), probs.data_ptr(), dlog.data_ptr() if want_grad else None,
                self._scratch.data_ptr(), self._scratch.numel() * 4, _lib.current_stream_ptr()))
        return self._scalars[0], probs, dlog

    def head_forward_backward(self, batch: int, targets: torch.Tensor):
        """Fused dense head: logits, probabilities, loss and the head's whole backward in one kernel (needs a
        preceding `forward(..., skip_head=True)`; `backward(x, None)` continues).  -> (loss, probs, logits)"""
        shape = self.cfg.target_shape(batch)
        targets = targets.contiguous()
        if tuple(targets.shape) != tuple(shape) or targets.dtype != torch.float32 or not targets.is_cuda:
            raise ValueError("targets must be CUDA float32 with the logits' shape")
        logits, probs = self._buf(("logits", batch), shape), self._buf(("probs", tuple(shape)), shape)
        ws = self._workspace(batch)
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_crnn_head_fwd_bwd(
                C.byref(self.desc), self.params.data_ptr(), batch, ws.data_ptr(), ws.numel(), targets.data_ptr(),
                self.loss_kind, self.alpha, self.gamma, 1.0, logits.data_ptr(), probs.data_ptr(),
                self._scalars.data_ptr(), self.grads.data_ptr(), _lib.current_stream_ptr()))
        return self._scalars[0], probs, logits

    def backward(self, x, dlogits, dx: torch.Tensor | None = None) -> torch.Tensor:
        """`dlogits=None`: continue after `head_forward_backward`."""
        x = self._check_x(x)
        B = x.shape[0]
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_crnn_backward(
                C.byref(self.desc), self.params.data_ptr(), x.data_ptr(), B, self._last_seed,
                ws.data_ptr(), ws.numel(), dlogits.data_ptr() if dlogits is not None else None, self.grads.data_ptr(),
                dx.data_ptr() if dx is not None else None, _lib.current_stream_ptr()))
        return self.grads

    def optimizer_step(self, world_size: int = 1) -> torch.Tensor:
        """clip (global norm) + Adam on the flat buffers; returns the pre-clip gradient norm (device scalar)."""
        self.step_count += 1
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_clip_adam(
                self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.params.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                self.step_count, self.clip, 1.0 / world_size, self._scalars[1:].data_ptr(),
                self._scratch.data_ptr(), self._scratch.numel() * 4, _lib.current_stream_ptr()))
        return self._scalars[1]

    def train_step(self, x: torch.Tensor, y: torch.Tensor):
        """One optimisation step on device-resident (x, y).  Returns (loss, probs) as device tensors
        (views of engine-owned buffers, valid until the next call)."""
        if self.fused_head and x.is_contiguous() and y.is_contiguous():
            from . import parallel
            single = self.xch is None and parallel.world_info(self.pg)[1] == 1
            if (single or self.xch is not None) and self.cuda_graph and self._ws is not None and self._ws_batch == x.shape[0]:
                out = self._train_step_graph(x, y)     # (the first call per batch size runs eagerly)
                if out is not None:
                    return out
            if single:
                return self._train_step_single(x, y)
        if self.xch is not None:
            self.grads = self.xch.next_grad_buffer()                  # this step's half of the exchange region
        if self.fused_head:
            self.forward(x, training=True, skip_head=True)
            loss, probs, _ = self.head_forward_backward(x.shape[0], y)
            dlog = None
        else:
            logits = self.forward(x, training=True)
            loss, probs, dlog = self.loss_and_grad(logits, y)
        if self.xch is not None:
            self.backward(x, dlog)
            self.step_count += 1
            self.grads = self.xch.allreduce_clip_adam(
                self.params, self.exp_avg, self.exp_avg_sq, step=self.step_count, lr=self.lr, betas=self.betas,
                eps=self.eps, weight_decay=self.weight_decay, clip=self.clip, gnorm_out=self._scalars[1:])
            return loss, probs
        self.backward(x, dlog)
        from . import parallel
        scale = parallel.allreduce_sum_(self.grads, self.pg)         # sum; 1/world folded into clip_adam
        world = round(1.0 / scale)
        self.optimizer_step(world)
        return loss, probs

    def _train_step_graph(self, x, y):
        """The single-process step as a CUDA graph: captured once per (x, y buffers, batch, hyper-parameters), then ONE
        launch per step.  Bit-identical to the eager step: the graph's first node advances the device-side step state
        to exactly the seed / bias corrections the eager path passes by value."""
        B = x.shape[0]
        xch = self.xch
        if xch is not None and xch.seq != self.step_count:
            return None                                # exchange number and optimizer step must move together
        parity = (self.step_count + 1) & 1 if xch is not None else 0
        gbuf = xch.grad_bufs[parity] if xch is not None else self.grads
        key = (x.data_ptr(), y.data_ptr(), B, self.lr, self.betas, self.eps, self.weight_decay, self.clip,
               float(self.desc.dropout), self.loss_kind, self.alpha, self.gamma, self.params.data_ptr(), gbuf.data_ptr())
        L, desc = self.L, C.byref(self.desc)
        if self._state_step != self.step_count:       # eager steps / reset_optimizer moved the host counter
            with torch.cuda.device(self.device):
                _lib.check(L.sedb200_step_state_init(self._step_state.data_ptr(), self.step_count, _lib.current_stream_ptr()))
            self._state_step = self.step_count
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 16 or self._graph_misses >= 64:
                return None                            # inputs arrive in ever-new buffers: stay eager
            self._graph_misses += 1
            x = self._check_x(x)
            shape = self.cfg.target_shape(B)
            if tuple(y.shape) != tuple(shape) or y.dtype != torch.float32 or not y.is_cuda:
                raise ValueError("targets must be CUDA float32 with the logits' shape")
            ws = self._workspace(B)
            logits, probs = self._buf(("logits", B), shape), self._buf(("probs", tuple(shape)), shape)
            p_state = self._step_state.data_ptr()
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            l0 = L.sedb200_launch_count()
            with torch.cuda.device(self.device), torch.cuda.graph(graph):
                st = _lib.current_stream_ptr()
                _lib.check(L.sedb200_step_advance(p_state, self.seed, self.betas[0], self.betas[1], st))
                _lib.check(L.sedb200_crnn_forward_s(desc, self.params.data_ptr(), self.bn_state.data_ptr(), x.data_ptr(), B, 1,
                                                    p_state, ws.data_ptr(), ws.numel(), None, st))
                _lib.check(L.sedb200_crnn_head_fwd_bwd(desc, self.params.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                       y.data_ptr(), self.loss_kind, self.alpha, self.gamma, 1.0,
                                                       logits.data_ptr(), probs.data_ptr(), self._scalars.data_ptr(),
                                                       gbuf.data_ptr(), st))
                _lib.check(L.sedb200_crnn_backward_s(desc, self.params.data_ptr(), x.data_ptr(), B, p_state, ws.data_ptr(),
                                                     ws.numel(), None, gbuf.data_ptr(), None, st))
                if xch is None:
                    _lib.check(L.sedb200_clip_adam_s(self.params.data_ptr(), gbuf.data_ptr(), self.exp_avg.data_ptr(),
                                                     self.exp_avg_sq.data_ptr(), self.params.numel(), self.lr, self.betas[0],
                                                     self.betas[1], self.eps, self.weight_decay, p_state, self.clip, 1.0,
                                                     self._scalars.data_ptr() + 4, self._scratch.data_ptr(),
                                                     self._scratch.numel() * 4, st))
                else:                                   # fused NVLink exchange + clip + Adam, this parity's buffers
                    _lib.check(L.sedb200_p2p_allreduce_clip_adam_s(
                        xch.table, xch.world, xch.rank, xch.n, parity, p_state, self.params.data_ptr(),
                        self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), xch.reduced.data_ptr(), self.lr, self.betas[0],
                        self.betas[1], self.eps, self.weight_decay, self.clip, 1.0 / xch.world,
                        self._scalars.data_ptr() + 4, xch.scratch.data_ptr(), xch.scratch.numel() * 4, st))
            self.launches_per_graph_step = int(L.sedb200_launch_count() - l0)
            entry = (graph, probs, ws, (x, y))            # the captured step reads these buffers: keep them alive
            self._graphs[key] = entry
        self._last_seed = self.seed + self.step_count
        if xch is not None:
            xch.raise_if_failed()
        entry[0].replay()
        if xch is not None:
            xch.after_graph_exchange()
            self.grads = xch.reduced
        self.step_count += 1
        self._state_step += 1
        self.num_batches_tracked += 1
        self.graph_replays += 1
        return self._scalars[0], entry[1]

    def _train_step_single(self, x, y):
        """train_step for the common single-process case (fused head, no gradient exchange): the same four library
        calls as the general route, with the argument marshalling done once -- a caller that reads the loss every
        step (sed.py:138) waits for this host code while the GPU is idle."""
        x = self._check_x(x)
        B = x.shape[0]
        shape = self.cfg.target_shape(B)
        if tuple(y.shape) != tuple(shape) or y.dtype != torch.float32 or not y.is_cuda:
            raise ValueError("targets must be CUDA float32 with the logits' shape")
        ws = self._workspace(B)
        logits, probs = self._buf(("logits", B), shape), self._buf(("probs", tuple(shape)), shape)
        L, desc = self.L, C.byref(self.desc)
        seed = self._last_seed = self.seed + self.step_count
        p_params, p_grads, p_ws, n_ws, p_x = (self.params.data_ptr(), self.grads.data_ptr(), ws.data_ptr(), ws.numel(),
                                              x.data_ptr())
        p_scal = self._scalars.data_ptr()
        with torch.cuda.device(self.device):
            st = _lib.current_stream_ptr()
            _lib.check(L.sedb200_crnn_forward(desc, p_params, self.bn_state.data_ptr(), p_x, B, 1, seed, p_ws, n_ws,
                                              None, st))
            self.num_batches_tracked += 1
            _lib.check(L.sedb200_crnn_head_fwd_bwd(desc, p_params, B, p_ws, n_ws, y.data_ptr(), self.loss_kind,
                                                   self.alpha, self.gamma, 1.0, logits.data_ptr(), probs.data_ptr(),
                                                   p_scal, p_grads, st))
            _lib.check(L.sedb200_crnn_backward(desc, p_params, p_x, B, seed, p_ws, n_ws, None, p_grads, None, st))
            self.step_count += 1
            _lib.check(L.sedb200_clip_adam(p_params, p_grads, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                           self.params.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                           self.weight_decay, self.step_count, self.clip, 1.0, p_scal + 4,
                                           self._scratch.data_ptr(), self._scratch.numel() * 4, st))
        return self._scalars[0], probs

    @torch.no_grad()
    def predict_proba(self, x: torch.Tensor, training_bn: bool = False) -> torch.Tensor:
        logits = self.forward(x, training=False) if not training_bn else self._forward_train_noupdate(x)
        return torch.sigmoid_(logits.clone())

    def _forward_train_noupdate(self, x):
        """train-mode BatchNorm forward that leaves running stats / counters untouched (parity checks)."""
        keep, nbt = self.bn_state.clone(), self.num_batches_tracked
        drop = self.desc.dropout
        self.desc.dropout = 0.0
        try:
            out = self.forward(x, training=True)
        finally:
            self.desc.dropout = drop
            self.bn_state.copy_(keep)
            self.num_batches_tracked = nbt
        return out

    def dropout_masks(self, batch: int, seed: int | None = None) -> list[torch.Tensor]:
        """Keep-masks (uint8 [B, conv_ch, H, W_out], one per conv block) of the training forward run with `seed`
        (default: the seed of the last forward).  The forward pass never stores them -- they come from the same
        counter-based generator -- so this is a test hook: an oracle can apply the SAME masks."""
        seed = self._last_seed if seed is None else int(seed)
        out, w = [], self.cfg.W
        for i, p in enumerate(self.cfg.pool):
            w //= p
            m = torch.empty(batch, self.cfg.conv_ch, self.cfg.H, w, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self.L.sedb200_crnn_dropout_mask(C.byref(self.desc), batch, seed, i, m.data_ptr(),
                                                            _lib.current_stream_ptr()))
            out.append(m)
        return out

    # ------------------------------------------------------------------ metrics on device
    def threshold_counts(self, probs: torch.Tensor, targets: torch.Tensor, block: int, threshold: float = 0.5):
        """13 integer counts behind metrics.py (see sedb200.h); probs/targets [..., n_cls] CUDA float32."""
        n_cls = probs.shape[-1]
        p2 = probs.reshape(-1, n_cls).contiguous()
        t2 = targets.reshape(-1, n_cls).contiguous().to(torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.L.sedb200_threshold_counts(p2.data_ptr(), t2.data_ptr(), p2.shape[0], n_cls, int(block),
                                                       float(threshold), self._counts.data_ptr(),
                                                       _lib.current_stream_ptr()))
        return self._counts
