"""ctypes binding of libvapb.so (include/vapb.h). There is no fallback: if the
library has not been built, importing the product path fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAPB_LIB") or os.path.join(_HERE, "libvapb.so")  # VAPB_LIB: A/B a second build

MODE_FP32, MODE_BF16, MODE_FP16, MODE_FP32_TC = 0, 1, 2, 3
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16, "fp16": MODE_FP16, "fp32_tc": MODE_FP32_TC}

# every symbol include/vapb.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
_fp = C.c_void_p  # device pointers travel as integers
SYMBOLS = {
    "vapb_create": (_i, [_i, C.POINTER(_vp)]),
    "vapb_load_tensor": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i64)]),
    "vapb_finalize": (_i, [_vp]),
    "vapb_destroy": (_i, [_vp]),
    "vapb_last_error": (C.c_char_p, [_vp]),
    "vapb_describe": (_i, [_vp] + [C.POINTER(_i)] * 5),
    "vapb_frames": (_i, [_i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "vapb_workspace_bytes": (_i, [_vp, _i, _i64, _i, C.POINTER(_sz)]),
    "vapb_forward": (_i, [_vp, _vp, _fp, _i, _i64, _i, _vp, _sz, _fp, _fp]),
    "vapb_forward_attention": (_i, [_vp, _vp, _fp, _i, _i64, _vp, _sz] + [_fp] * 5),
    "vapb_probs": (_i, [_vp, _vp, _fp, _i, _i64, _i, _vp, _sz, _i, _i, _i, _i] + [_fp] * 9),
    "vapb_probs_ex": (_i, [_vp, _vp, _fp, _i, _i, _i64, _i, _vp, _sz, _i, _i, _i, _i] + [_fp] * 10),
    "vapb_pcm16_to_f32": (_i, [_vp, _fp, _i64, _fp]),
    "vapb_memset_zero": (_i, [_vp, _fp, _sz]),
    "vapb_probs_from_logits": (_i, [_vp, _vp, _fp, _i64, _i, _i, _i, _i] + [_fp] * 5),
    "vapb_get_stage": (_i, [_vp, _vp, C.c_char_p, _i, _i64, _i, _vp, _sz, _fp, _sz]),
    "vapb_profile_begin": (_i, [_vp]),
    "vapb_profile_end": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "vapb_debug_gemm_x3": (_i, [_vp, _fp, _i64, _i64, _vp, _i, _i, _i, _i, _fp, _i, _fp, _fp, _i, _fp, _i, _fp, _i,
                                _fp, _fp, _fp, C.c_char_p, _i]),
    "vapb_debug_gemm_2sm": (_i, [_vp, _fp, _i64, _i64, _fp, _i, _i, _i, _fp, _i, _fp, _fp, _i, _fp, C.c_char_p, _i]),
    "vapb_debug_conv0_tc": (_i, [_vp, _fp, _i, _i64, _vp, _vp, _vp, _vp, _fp, _i, C.c_char_p, _i]),
    "vapb_debug_conv01": (_i, [_vp, _fp, _i, _i64, _vp, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _fp, _i64, _i, _i, C.c_char_p, _i, _fp]),
    "vapb_debug_gemm_lin": (_i, [_vp, _fp, _i64, _i64, _fp, _i, _i, _i, _i, _fp, _i, _fp, _fp, _i, _fp, _i,
                                 _fp, _i, _fp, _i, _fp, _fp, _fp, C.c_char_p, _i]),
    "vapb_debug_ffn_fused": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, C.c_char_p, _i, _fp]),
    "vapb_debug_rnn_pack": (_i, [_i, _fp, _fp, _fp, _fp, _fp, _fp]),
    "vapb_debug_rnn_tc": (_i, [_vp, _i, _fp, _i64, _i64, _fp, _fp, _fp, _i64, _i, _i, C.c_char_p, _i, _fp, _i]),
    "vapb_debug_attn_tc": (_i, [_vp, _fp, _i64, _fp, _fp, _i64, _fp, _i, _i, _i, _fp, _i, C.c_char_p, _i, _fp]),
    "vapb_debug_attn_x3": (_i, [_vp, _fp, _i, _fp, _i, _i, _i, _fp, _fp, _i, _i, _fp, _i, C.c_char_p, _i]),
    "vapb_vad_filter": (_i, [_vp, _vp, _fp, _i, _i64, _i, _i, _fp]),
    "vapb_vad_filter_ex": (_i, [_vp, _vp, _fp, _i, C.c_float, _i, _i64, _i, _i, _fp]),
    "vapb_zero_shot": (_i, [_vp, _vp, _fp, _i, _i64, _i64, _fp, _i64, _vp, _fp, _fp, _fp, _fp]),
    "vapb_resample": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i64, _i64, _i64, _i64, _i, _i, _i, _fp, _fp, _i64, _i64]),
    "vapb_launch_count": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "vapb_build_info": (C.c_char_p, []),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -m voiceactivityprojection_b200.build). There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


PROFILE_FAMILIES = ["conv0", "conv_gemm", "linear_gemm", "attention", "rnn", "heads", "other"]


class VapbError(RuntimeError):
    pass


def check(lib, handle, rc):
    if rc != 0:
        msg = lib.vapb_last_error(handle)
        raise VapbError(f"vapb error {rc}: {msg.decode() if msg else ''}")


def frames(n_samples: int):
    """(frames @100 Hz, frames @50 Hz) for n_samples (vapb_frames)."""
    lib = load()
    a, b = _i64(), _i64()
    rc = lib.vapb_frames(n_samples, C.byref(a), C.byref(b))
    if rc != 0:
        raise VapbError(f"n_samples={n_samples} is too short for the encoder")
    return a.value, b.value
