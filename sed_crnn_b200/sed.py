"""Drop-in for the model + epoch runner of /root/reference/sed.py (sed.py:82-141).

`TimePooledCRNN(conv_channels=128, dropout=0.5)` keeps the reference's module names / state_dict keys;
`run_epoch(model, loader, loss_fn, optim=None)` is the reference's loop (sed.py:128-141) verbatim in
behaviour, including the per-step loss read and prediction collection."""
from __future__ import annotations

import numpy as np
import torch

from .decorte_datamodule import HitWindowDataset, _find_clean_negatives as find_clean_negatives   # noqa: F401  (sed.py:48-79)
from .decorte_datamodule import _load_all_npz as load_all_npz                  # noqa: F401  (sed.py:115-125)
from .modules import BCEWithLogitsLoss, SedTimePooledCRNN as TimePooledCRNN   # noqa: F401
from .train_constants import FPS_OUT, SEQ_LEN_IN, SEQ_LEN_OUT, TIME_POOL       # noqa: F401

DEVICE = torch.device("cuda")


def run_epoch(model, loader, loss_fn, optim=None):
    train = optim is not None
    model.train() if train else model.eval()
    total, preds, labels = 0., [], []
    for xb, yb in loader:
        xb, yb = xb.to(DEVICE), yb.to(DEVICE)
        if train:
            optim.zero_grad()
        with torch.set_grad_enabled(train):
            out = model(xb)
            loss = loss_fn(out, yb)
        if train:
            loss.backward()
            optim.step()
        total += loss.item()
        preds.append(torch.sigmoid(out).detach().cpu().numpy())
        labels.append(yb.detach().cpu().numpy())
    return total / len(loader), np.concatenate(preds), np.concatenate(labels)
