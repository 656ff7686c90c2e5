"""ctypes binding of libsedb200.so (include/sedb200.h).  No CPU fallback: if the library is missing or
the device is not a B200-class GPU every compute call raises."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsedb200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "sedb200.h")

OK, EINVAL, ESHAPE, EWORKSPACE, ECUDA, EARCH = 0, -1, -2, -3, -4, -5
PAD_MODES = {"constant": 0, "reflect": 1}


class Sedb200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsedb200 error {code}: {msg}")
        self.code = code


_lib = None

_p, _i, _l, _f, _sz = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_size_t
# name -> (restype, argtypes); must list every symbol include/sedb200.h declares
SIGNATURES = {
    "sedb200_version": (_i, []),
    "sedb200_last_error": (C.c_char_p, []),
    "sedb200_device_check": (_i, [_i]),
    "sedb200_logmel_frames": (_l, [_l]),
    "sedb200_logmel_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
    "sedb200_logmel_host_scratch": (_sz, [_i, _i, _l]),
    "sedb200_logmel_host_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _sz, _p]),
    "sedb200_mel_filterbank": (_i, [_i, _p]),
}


def header_symbols() -> list[str]:
    """Function names declared in include/sedb200.h."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(sedb200_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Sedb200Error(EARCH, f"{LIB_PATH} not built (run `python -m sed_crnn_b200.build`); "
                                      "there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise Sedb200Error(rc, lib().sedb200_last_error().decode(errors="replace"))


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
