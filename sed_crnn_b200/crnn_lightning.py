"""Drop-in for /root/reference/crnn_lightning.py: FocalBCELoss, TimePooledCRNN, CRNNLightning.

Same public names, constructor signatures, attributes (`model`, `loss_fn`, `hparams`, `_buf`, `track`) and
Lightning hooks as the reference (crnn_lightning.py:79-200); the arithmetic runs on libsedb200.so.  The
epoch-end aggregation (crnn_lightning.py:102-129) thresholds and counts ON THE DEVICE instead of moving
every prediction of the epoch to the host; the scores are bit-identical to metrics.py for identical
decisions.  Plotting (crnn_lightning.py:131-154) is out of scope and skipped when matplotlib is absent.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import metrics
from .modules import FocalBCELoss, TimePooledCRNN          # noqa: F401  (re-exported, reference names)
from .train_constants import FPS_OUT

EPS = 1e-12          # crnn_lightning.py:21

try:                                                       # real Lightning when it is installed
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                                          # pragma: no cover - not installed in this image
    class _HParams(dict):
        __getattr__ = dict.__getitem__

    class _Base(nn.Module):
        """Just enough of LightningModule for the hooks below to run under a hand-written loop."""

        def __init__(self):
            super().__init__()
            self.hparams = _HParams()
            self.current_epoch = 0
            self.logged = {}

        def save_hyperparameters(self, *args, ignore=()):
            import inspect
            loc = inspect.currentframe().f_back.f_locals
            for k, v in loc.items():
                if k not in ("self", "__class__") and k not in ignore:
                    self.hparams[k] = v

        def log(self, name, value, **kw):
            self.logged[name] = value


class CRNNLightning(_Base):
    def __init__(self, fold_id: int, art_dir: str, lr=1e-3, weight_decay=1e-4, dropout=0.4):
        super().__init__()
        self.save_hyperparameters(ignore=["art_dir"])
        self.art_dir = art_dir
        self.model = TimePooledCRNN(dropout)
        self.loss_fn = FocalBCELoss()
        self._buf = {m: {"preds": [], "trues": [], "losses": []} for m in ["train", "val"]}
        self.track = {k: [] for k in [
            "loss_tr", "loss_val", "f1_1s_tr", "f1_1s_val", "er_1s_tr", "er_1s_val",
            "f1_fr_tr", "f1_fr_val", "er_fr_tr", "er_fr_val"]}

    def forward(self, x):
        return self.model(x)

    # ---- helpers (crnn_lightning.py:97-129)
    def _collect(self, logits, y, loss, mode):
        self._buf[mode]["preds"].append(torch.sigmoid(logits.detach()))
        self._buf[mode]["trues"].append(y)
        self._buf[mode]["losses"].append(loss.detach())

    def _aggregate(self, mode):
        p_t = torch.cat(self._buf[mode]["preds"])
        t_t = torch.cat(self._buf[mode]["trues"])
        loss = torch.stack(self._buf[mode]["losses"]).mean().item()
        step_rows = [p.shape[0] * p.shape[1] for p in self._buf[mode]["preds"]]
        for k in self._buf[mode]:
            self._buf[mode][k].clear()
        from . import parallel
        if parallel.world_info()[1] > 1:
            # data-parallel epoch: decisions are re-assembled in single-process order, counted on block-aligned shards
            # and summed with one integer all-reduce (parallel.sharded_metric_counts) -- same scores as one process
            n_cls = p_t.shape[-1]
            c = parallel.sharded_metric_counts((p_t > 0.5).reshape(-1, n_cls), t_t.reshape(-1, n_cls), step_rows,
                                               FPS_OUT)
            total_t = torch.tensor([p_t.numel()], dtype=torch.int64, device=p_t.device)
            total = int(parallel.allreduce_counts_(total_t)[0])
        else:
            c = metrics._counts(p_t, t_t, FPS_OUT)                   # 13 integers, one small D2H
            total = p_t.numel()
        f1_fr, er_fr, f1_1s, er_1s = metrics.scores_from_counts(c)
        tp, nsys, nref = int(c[0]), int(c[1]), int(c[2])
        fp, fn = nsys - tp, nref - tp
        cm = np.array([[total - tp - fp - fn, fp], [fn, tp]])
        return dict(loss=loss, f1_frame=f1_fr, er_frame=er_fr, f1_1s=f1_1s, er_1s=er_1s, cm=cm)

    def _plot_epoch(self, epoch, tr, val):
        return None                                                  # visualisation: out of scope

    # ---- Lightning hooks (crnn_lightning.py:157-200)
    def training_step(self, batch, _):
        x, y = batch
        logits = self(x)
        loss = self.loss_fn(logits, y)
        self._collect(logits, y, loss, "train")
        self.log("train_loss", loss, on_epoch=True, prog_bar=True)
        return loss

    def on_train_epoch_end(self):
        tr = self._aggregate("train")
        self.track["loss_tr"].append(tr["loss"])
        self.track["f1_1s_tr"].append(tr["f1_1s"]); self.track["er_1s_tr"].append(tr["er_1s"])
        self.track["f1_fr_tr"].append(tr["f1_frame"]); self.track["er_fr_tr"].append(tr["er_frame"])
        self._last_train = tr

    def validation_step(self, batch, _):
        x, y = batch
        logits = self(x)
        loss = self.loss_fn(logits, y)
        self._collect(logits, y, loss, "val")
        self.log("val_loss", loss, on_epoch=True, prog_bar=True)

    def on_validation_epoch_end(self):
        val = self._aggregate("val")
        self.track["loss_val"].append(val["loss"])
        self.track["f1_1s_val"].append(val["f1_1s"]); self.track["er_1s_val"].append(val["er_1s"])
        self.track["f1_fr_val"].append(val["f1_frame"]); self.track["er_fr_val"].append(val["er_frame"])
        self.log("val_er_1s", val["er_1s"], prog_bar=True)
        self.log("val_f1_1s", val["f1_1s"], prog_bar=True)
        if not hasattr(self, "_last_train"):
            self._last_train = val.copy()
        self._plot_epoch(self.current_epoch, self._last_train, val)

    def configure_optimizers(self):
        opt = torch.optim.Adam(self.parameters(), lr=self.hparams.lr, weight_decay=self.hparams.weight_decay)
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=.5, patience=10)
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "monitor": "val_loss"}}
