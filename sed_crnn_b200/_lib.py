"""ctypes binding of libsedb200.so (include/sedb200.h).  No CPU fallback: if the library is missing or
the device is not a B200-class GPU every compute call raises."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEDB200_LIB_PATH") or os.path.join(HERE, "libsedb200.so")   # override: A/B builds
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
    "sedb200_launch_count": (_l, []),
    "sedb200_prof_enable": (_i, [_i]),
    "sedb200_prof_report": (_i, [C.c_char_p, _sz]),
    "sedb200_logmel_frames": (_l, [_l]),
    "sedb200_logmel_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
    "sedb200_logmel_host_scratch": (_sz, [_i, _i, _l]),
    "sedb200_logmel_host_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _sz, _p]),
    "sedb200_logmel_i16": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
    "sedb200_logmel_host_scratch_i16": (_sz, [_i, _i, _l]),
    "sedb200_logmel_host_i16": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _sz, _p]),
    "sedb200_logmel_f32_k": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _i]),
    "sedb200_logmel_i16_k": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _i]),
    "sedb200_logmel_default_kernel": (_i, []),
    "sedb200_mel_filterbank": (_i, [_i, _p]),
    "sedb200_standardize_scratch_bytes": (_sz, [_l, _i]),
    "sedb200_standardize_fit": (_i, [_p, _l, _i, _p, _p, _p, _p, _sz, _p]),
    "sedb200_standardize_apply": (_i, [_p, _l, _i, _p, _p, _p, _p]),
    "sedb200_crnn_validate": (_i, [_p]),
    "sedb200_crnn_seq_len": (_i, [_p]),
    "sedb200_crnn_flat": (_i, [_p]),
    "sedb200_crnn_n_tensors": (_i, [_p]),
    "sedb200_crnn_param_layout": (_l, [_p, _p]),
    "sedb200_crnn_bn_state_floats": (_l, [_p]),
    "sedb200_crnn_workspace_bytes": (_sz, [_p, _i]),
    "sedb200_crnn_forward": (_i, [_p, _p, _p, _p, _i, _i, C.c_ulonglong, _p, _sz, _p, _p]),
    "sedb200_loss_fwd_bwd": (_i, [_i, _f, _f, _p, _p, _l, _f, _p, _p, _p, _p, _sz, _p]),
    "sedb200_loss_scratch_bytes": (_sz, [_l]),
    "sedb200_crnn_head_supported": (_i, [_p]),
    "sedb200_crnn_head_fwd_bwd": (_i, [_p, _p, _i, _p, _sz, _p, _i, _f, _f, _f, _p, _p, _p, _p, _p]),
    "sedb200_crnn_backward": (_i, [_p, _p, _p, _i, C.c_ulonglong, _p, _sz, _p, _p, _p, _p]),
    "sedb200_crnn_dropout_mask": (_i, [_p, _i, C.c_ulonglong, _i, _p, _p]),
    "sedb200_gru_scan_fused_bias_grads": (_i, [_i]),
    "sedb200_gru_scan_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "sedb200_gru_scan_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "sedb200_clip_adam_scratch_bytes": (_sz, [_l]),
    "sedb200_clip_adam": (_i, [_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _l, _f, _f, _p, _p, _sz, _p]),
    "sedb200_step_state_bytes": (_sz, []),
    "sedb200_step_state_init": (_i, [_p, _l, _p]),
    "sedb200_step_advance": (_i, [_p, C.c_ulonglong, _f, _f, _p]),
    "sedb200_crnn_forward_s": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _sz, _p, _p]),
    "sedb200_crnn_backward_s": (_i, [_p, _p, _p, _i, _p, _p, _sz, _p, _p, _p, _p]),
    "sedb200_clip_adam_s": (_i, [_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _p, _f, _f, _p, _p, _sz, _p]),
    "sedb200_p2p_region_bytes": (_sz, [_l]),
    "sedb200_p2p_grad_offset_bytes": (_l, [_l, _i]),
    "sedb200_p2p_region_alloc": (_i, [_sz, _p, _p]),
    "sedb200_p2p_region_open": (_i, [_p, _p]),
    "sedb200_p2p_region_close": (_i, [_p]),
    "sedb200_p2p_region_free": (_i, [_p]),
    "sedb200_p2p_status": (_i, [_p, _p]),
    "sedb200_p2p_scratch_bytes": (_sz, []),
    "sedb200_p2p_status_offset_bytes": (_l, []),
    "sedb200_p2p_allreduce_clip_adam": (_i, [_p, _i, _i, _l, _l, _l, _p, _p, _p, _p, _f, _f, _f, _f, _f, _f, _f, _p,
                                            _p, _sz, _p]),
    "sedb200_p2p_allreduce_clip_adam_s": (_i, [_p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _f, _f, _f, _f, _f, _f, _f, _p,
                                              _p, _sz, _p]),
    "sedb200_threshold_counts": (_i, [_p, _p, _l, _i, _i, _f, _p, _p]),
    "sedb200_window_batch_f32": (_i, [_p, _p, _l, _i, _i, _i, _p, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "sedb200_clean_negatives": (_i, [_p, _l, _i, _i, _p, _p]),
    "sedb200_rasterize_labels": (_i, [_p, _p, _i, _i, _i, _l, _i, _i, _p, _p]),
    "sedb200_conv3x3_tc_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sedb200_conv3x3_wgrad_tc_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sedb200_conv3x3_wgrad_tc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sedb200_conv3x3_planes_test_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sedb200_conv3x3_planes_test": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sedb200_conv3x3_small_supported": (_i, [_i, _i, _i, _i]),
    "sedb200_conv3x3_small": (_i, [_p, _l, _l, _l, _l, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    "sedb200_conv3x3_small_wgrad_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "sedb200_conv3x3_small_wgrad": (_i, [_p, _p, _l, _l, _l, _l, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "sedb200_gemm_tc_scratch_bytes": (_sz, [_i, _i, _i]),
    "sedb200_gemm_tc": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _p, _sz, _p]),
    "sedb200_conv3x3_tc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
}


class CrnnDesc(C.Structure):
    """mirror of `sedb200_crnn_desc` (include/sedb200.h)"""
    _fields_ = [("mode", _i), ("in_ch", _i), ("H", _i), ("W", _i), ("n_conv", _i), ("conv_ch", _i),
                ("pool", _i * 4), ("n_gru", _i), ("gru_units", _i * 4), ("n_dense", _i),
                ("dense_units", _i * 3), ("dense_relu", _i), ("dropout", _f), ("dropout_each_block", _i),
                ("bn_eps", _f), ("bn_momentum", _f), ("tensor_cores", _i)]


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
