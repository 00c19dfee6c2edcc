"""ctypes binding of libb200codec.so (the C ABI declared in include/b200codec.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C tts_max_b200/csrc`.
Loading fails loudly when it is missing: this package never falls back to another backend.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

ABI_VERSION = 3
PRECISIONS = {"bf16": 0, "fp16": 1}
IDS_I32, IDS_I64 = 0, 1
DT_F32, DT_F16, DT_BF16, DT_F64 = 0, 1, 2, 3

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200codec.so")


class B200CodecConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", c_int32),
        ("sample_rate", c_int32),
        ("hop_length", c_int32),
        ("n_upsample", c_int32),
        ("precision", c_int32),
        ("device", c_int32),
        ("hidden_dim", c_int32),
        ("depth", c_int32),
        ("heads", c_int32),
        ("vq_dim", c_int32),
        ("upsample_factors", c_int32 * 3),
        ("kernel_sizes", c_int32 * 3),
    ]


# name -> (restype, argtypes); one entry per function declared in include/b200codec.h
SIGNATURES = {
    "b200codec_last_error": (c_char_p, []),
    "b200codec_num_tensors": (c_int, [c_void_p]),
    "b200codec_tensor_key": (c_char_p, [c_void_p, c_int]),
    "b200codec_tensor_shape": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "b200codec_create": (c_int, [POINTER(B200CodecConfig), POINTER(c_void_p)]),
    "b200codec_destroy": (None, [c_void_p]),
    "b200codec_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64), c_int]),
    "b200codec_read_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t]),
    "b200codec_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "b200codec_decode_varlen": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p]),
    "b200codec_decode_host": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p]),
    "b200codec_decode_host_async": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p]),
    "b200codec_take_id_error": (c_int, [c_void_p]),
    "b200codec_stream_create": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
    "b200codec_stream_destroy": (None, [c_void_p]),
    "b200codec_stream_reset": (c_int, [c_void_p]),
    "b200codec_stream_capacity": (c_int, [c_void_p]),
    "b200codec_stream_tokens": (c_int64, [c_void_p]),
    "b200codec_stream_push": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b200codec_plan_generation": (c_int64, [c_void_p]),
    "b200codec_set_stage_taps": (c_int, [c_void_p, c_int]),
    "b200codec_stage_width": (c_int, [c_void_p, c_char_p]),
    "b200codec_stage_rows": (c_int64, [c_void_p, c_char_p]),
    "b200codec_read_stage": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t, c_void_p]),
    "b200codec_set_zero_copy_output": (c_int, [c_int]),
    "b200codec_set_gemm_narrow_tiles": (c_int, [c_int]),
    "b200codec_set_istft_tile": (c_int, [c_int]),
    "b200codec_set_gemm_chain": (c_int, [c_int]),
    "b200codec_set_gemm_early_weights": (c_int, [c_int]),
    "b200codec_set_frontend_fold": (c_int, [c_int]),
    "b200codec_set_pdl": (c_int, [c_int]),
    "b200codec_samples_per_token": (c_int, [c_void_p]),
    "b200codec_launch_count": (c_int64, [c_void_p]),
    "b200codec_profile": (c_int, [c_void_p, c_int]),
    "b200codec_stage_times": (c_int, [c_void_p, c_int, POINTER(c_char_p), POINTER(c_float), POINTER(c_int)]),
    "b200codec_fsq_lookup": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "b200codec_map_speech_tokens": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200codec_fsq_quantize": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int,
                                       c_void_p]),
    "b200codec_istft": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p]),
    "b200codec_gemm": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "b200codec_rmsnorm": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "b200codec_layernorm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "b200codec_groupnorm_swish": (c_int, [c_int, c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_int, c_int,
                                          c_float, c_void_p, c_void_p]),
    "b200codec_attention": (c_int, [c_int, c_void_p, POINTER(c_int32), c_int, c_int, c_void_p, c_void_p]),
    # encode direction (b200enc_*)
    "b200enc_create": (c_int, [c_int, c_int, POINTER(c_void_p)]),
    "b200enc_destroy": (None, [c_void_p]),
    "b200enc_num_tensors": (c_int, [c_void_p]),
    "b200enc_tensor_key": (c_char_p, [c_void_p, c_int]),
    "b200enc_tensor_shape": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "b200enc_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "b200enc_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "b200enc_encode": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "b200enc_encode_batch": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "b200enc_set_stage_taps": (c_int, [c_void_p, c_int]),
    "b200enc_read_stage": (c_int, [c_void_p, c_char_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "b200enc_launch_count": (c_int64, [c_void_p]),
}

_lib = None


class B200CodecError(RuntimeError):
    """A C-ABI call returned non-zero."""


def load() -> ctypes.CDLL:
    """Loads libb200codec.so (once) and binds every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(`python -c 'import __graft_entry__ as g; g.build()'` or `make -C tts_max_b200/csrc`). "
            "tts_max_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


_VALUE_ERROR_PREFIXES = ("speech id", "decode: utterance", "decode: empty", "Current hop length")


def check(rc: int) -> None:
    """Turns a non-zero status into the exception class the reference would raise."""
    if rc == 0:
        return
    msg = load().b200codec_last_error().decode("utf-8", "replace")
    if msg.startswith(_VALUE_ERROR_PREFIXES):
        raise ValueError(msg)
    if "not supported yet" in msg:
        raise NotImplementedError(msg)
    raise B200CodecError(msg)


def i32_array(values) -> ctypes.Array:
    arr = (c_int32 * len(values))(*[int(v) for v in values])
    return arr
