"""CRNN hyper-parameters as a frozen dataclass.

The reference hard-codes these as module constants (/root/reference/train_constants.py:6-28,
sed.py:24-36); the defaults below ARE those constants, and the presets add the BASELINE.json configs.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field, replace

from . import _lib


@dataclass(frozen=True)
class CRNNConfig:
    mode: str = "fork"                     # "fork": input [B,Cin,F,T], pool over time (crnn_lightning.py:46-50)
    #                                        "sednet": input [B,Cin,T,F], pool over mel (BASELINE configs)
    in_ch: int = 1
    n_freq: int = 40                       # train_constants.N_MELS
    seq_len: int = 64                      # train_constants.SEQ_LEN_IN
    conv_ch: int = 16                      # train_constants.CONV_DEPTH
    pool: tuple = (2, 2, 2)                # train_constants.TIME_POOL
    gru_units: tuple = (16, 8)             # GRU1_UNITS, GRU2_UNITS
    dense_units: tuple = (8,)              # DENSE1_UNITS (hidden dense layers)
    n_classes: int = 1
    dense_relu: bool = True                # crnn_lightning.py:72
    dropout: float = 0.4                   # crnn_lightning.py:42
    dropout_each_block: bool = False       # sed.py:107 -> True
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    tensor_cores: bool = True              # conv contractions on tcgen05 (3-term bf16 split) where the shape allows

    # ---- geometry
    @property
    def H(self) -> int:
        return self.n_freq if self.mode == "fork" else self.seq_len

    @property
    def W(self) -> int:
        return self.seq_len if self.mode == "fork" else self.n_freq

    @property
    def seq_len_out(self) -> int:
        return self.seq_len // math.prod(self.pool) if self.mode == "fork" else self.seq_len

    @property
    def flat(self) -> int:
        w = self.W
        for p in self.pool:
            w //= p
        return self.conv_ch * (self.H if self.mode == "fork" else w)

    def input_shape(self, batch: int) -> tuple:
        return (batch, self.in_ch, self.H, self.W)

    def target_shape(self, batch: int) -> tuple:
        return (batch, self.seq_len_out, self.n_classes)

    def desc(self) -> _lib.CrnnDesc:
        d = _lib.CrnnDesc()
        d.mode = {"fork": 0, "sednet": 1}[self.mode]
        d.in_ch, d.H, d.W = self.in_ch, self.H, self.W
        d.n_conv, d.conv_ch = len(self.pool), self.conv_ch
        for i, p in enumerate(self.pool):
            d.pool[i] = p
        d.n_gru = len(self.gru_units)
        for i, h in enumerate(self.gru_units):
            d.gru_units[i] = h
        dense = list(self.dense_units) + [self.n_classes]
        d.n_dense = len(dense)
        for i, u in enumerate(dense):
            d.dense_units[i] = u
        d.dense_relu = int(self.dense_relu)
        d.dropout = float(self.dropout)
        d.dropout_each_block = int(self.dropout_each_block)
        d.bn_eps, d.bn_momentum = float(self.bn_eps), float(self.bn_momentum)
        d.tensor_cores = int(self.tensor_cores)
        return d

    # ---- flat-parameter layout (canonical tensor names, shapes and offsets)
    def tensor_specs(self) -> list[tuple[str, tuple, int]]:
        """[(name, shape, offset_in_floats)] in the C library's canonical order (sedb200.h)."""
        d = self.desc()
        L = _lib.lib()
        nt = L.sedb200_crnn_n_tensors(C.byref(d))
        if nt < 0:
            _lib.check(L.sedb200_crnn_validate(C.byref(d)))
        offs = (C.c_long * nt)()
        total = L.sedb200_crnn_param_layout(C.byref(d), offs)
        names: list[tuple[str, tuple]] = []
        cin = self.in_ch
        for i in range(len(self.pool)):
            names += [(f"conv{i}.weight", (self.conv_ch, cin, 3, 3)), (f"conv{i}.bias", (self.conv_ch,)),
                      (f"bn{i}.weight", (self.conv_ch,)), (f"bn{i}.bias", (self.conv_ch,))]
            cin = self.conv_ch
        gin = self.flat
        for i, h in enumerate(self.gru_units):
            names += [(f"gru{i}.w_ih", (2, 3 * h, gin)), (f"gru{i}.w_hh", (2, 3 * h, h)),
                      (f"gru{i}.b_ih", (2, 3 * h)), (f"gru{i}.b_hh", (2, 3 * h))]
            gin = 2 * h
        for i, u in enumerate(list(self.dense_units) + [self.n_classes]):
            names += [(f"dense{i}.weight", (u, gin)), (f"dense{i}.bias", (u,))]
            gin = u
        assert len(names) == nt
        del total
        return [(n, s, int(offs[k])) for k, (n, s) in enumerate(names)]

    def n_param_floats(self) -> int:
        d = self.desc()
        return int(_lib.lib().sedb200_crnn_param_layout(C.byref(d), None))


# the tree as shipped (train_constants.py) -- crnn_lightning.TimePooledCRNN
FORK = CRNNConfig()
# sed.py:82-112 -- 128 channels, dropout .5 after every block, nn.GRU(num_layers=2, hidden 32), fc 64->1
SEDPY = CRNNConfig(conv_ch=128, gru_units=(32, 32), dense_units=(), dropout=0.5, dropout_each_block=True)
# BASELINE.json configs[0], [1], [4]
C1 = CRNNConfig(mode="sednet", in_ch=1, seq_len=256, conv_ch=128, pool=(5, 2, 2), gru_units=(32, 32),
                dense_units=(16,), n_classes=6, dropout=0.5, dropout_each_block=True)
C2 = replace(C1, in_ch=2)
C5 = CRNNConfig(mode="sednet", in_ch=1, seq_len=2048, conv_ch=256, pool=(5, 2, 2), gru_units=(128, 128, 128),
                dense_units=(16,), n_classes=16, dropout=0.5, dropout_each_block=True)

PRESETS = {"fork": FORK, "sedpy": SEDPY, "c1": C1, "c2": C2, "c5": C5}
