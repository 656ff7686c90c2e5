"""Import the reference's own Python files, unmodified, from /root/reference.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU box, so
nothing that runs there (tests -m gpu, smoke(), bench.py) may call this.  It is used by
oracle/make_golden.py (which writes tests/golden/*.npz) and by the `not gpu` tests that pin the
restated oracles against the real reference when the tree is present.

crnn_lightning.py / sed.py import `matplotlib` and `pytorch_lightning`, neither of which is
installed and neither of which does arithmetic; both are stubbed.  The reference modules also
`os.makedirs` under `~` at import time (sed.py:41, feature.py:34, train_lightning.py:21), so
HOME is pointed at a throw-away directory first.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("SED_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "crnn_lightning.py"))


def _install_stubs() -> None:
    import torch.nn as nn

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None

        def _plt_getattr(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _noop

        plt.__getattr__ = _plt_getattr                # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class _HParams(dict):
            __getattr__ = dict.__getitem__

        class LightningModule(nn.Module):
            def __init__(self):
                super().__init__()
                self.hparams = _HParams()
                self.current_epoch = 0
                self.logged = {}

            def save_hyperparameters(self, *names, ignore=(), frame_locals=None):
                import inspect
                fr = inspect.currentframe().f_back
                loc = dict(fr.f_locals)
                for k, v in loc.items():
                    if k in ("self", "__class__") or k in ignore:
                        continue
                    self.hparams[k] = v

            def log(self, name, value, **kw):
                self.logged[name] = value

        class LightningDataModule:
            def __init__(self):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules["pytorch_lightning"] = pl


_loaded: dict[str, types.ModuleType] = {}


def load(name: str) -> types.ModuleType:
    """Return the reference module `name` (e.g. 'metrics', 'crnn_lightning', 'sed')."""
    if name in _loaded:
        return _loaded[name]
    if not available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} not present (expected on the GPU box)")
    _install_stubs()
    old_home = os.environ.get("HOME")
    tmp_home = tempfile.mkdtemp(prefix="sedref_home_")
    os.environ["HOME"] = tmp_home
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        # the reference's `utils` / `metrics` are generic names: import under those names
        # (its own files import each other that way) but keep them out of the way afterwards.
        saved = {k: sys.modules.pop(k) for k in ("utils", "metrics", "train_constants") if k in sys.modules
                 and not getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT)}
        mod = importlib.import_module(name)
        _loaded[name] = mod
        for k in ("utils", "metrics", "train_constants", "crnn_lightning", "sed"):
            if k in sys.modules and getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT):
                _loaded.setdefault(k, sys.modules[k])
        sys.modules.update(saved)
        return mod
    finally:
        sys.path.remove(REFERENCE_ROOT)
        if old_home is None:
            os.environ.pop("HOME", None)
        else:
            os.environ["HOME"] = old_home
