"""ctypes binding of libmsclip_b200.so (the C ABI of include/msclip_b200.h and msclip_b200_ops.h).

There is no Python / CPU fallback: if the shared library is missing this module raises, and every
compute entry point of the library fails on a box without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsclip_b200.so")

F32, BF16, F16, I64 = 0, 1, 2, 3
EPI_BF16, EPI_QGELU_BF16, EPI_RELU_BF16, EPI_RESID_F32, EPI_F32 = range(5)


class MsclipError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("patch_size", C.c_int32), ("layers", C.c_int32), ("width", C.c_int32), ("embed_dim", C.c_int32),
        ("image_resolution", C.c_int32), ("context_length", C.c_int32), ("vocab_size", C.c_int32),
        ("early_strides", C.c_int32 * 4), ("parallel_strides", C.c_int32 * 5), ("t2b_kernels", C.c_int32 * 5),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGNATURES = {
    # include/msclip_b200.h
    "msclip_version": (C.c_char_p, []),
    "msclip_last_error": (C.c_char_p, []),
    "msclip_device_count": (_I, []),
    "msclip_create": (_I, [C.POINTER(Config), C.POINTER(_P)]),
    "msclip_destroy": (_I, [_P]),
    "msclip_set_weight": (_I, [_P, C.c_char_p, _P, _I, _I, C.POINTER(_L)]),
    "msclip_finalize_weights": (_I, [_P, _P]),
    "msclip_logit_scale_exp": (_I, [_P, C.POINTER(_F)]),
    "msclip_encode_image": (_I, [_P, _P, _I, _I, _P, _I, _P]),
    "msclip_encode_text": (_I, [_P, _P, _I, _P, _I, _P]),
    "msclip_similarity_logits": (_I, [_P, _P, _I, _P, _I, _F, _P, _P]),
    "msclip_forward": (_I, [_P, _P, _I, _P, _I, _P, _P]),
    "msclip_comm_init": (_I, [_P, _I, _I, _I]),
    "msclip_comm_export": (_I, [_P, _P]),
    "msclip_comm_import": (_I, [_P, _P]),
    "msclip_contrastive_loss": (_I, [_P, _I, _F, _P, _P, _P]),
    "msclip_forward_loss": (_I, [_P, _P, _I, _P, _I, _P, _P, _P]),
    "msclip_launch_count": (_L, [_P]),
    "msclip_device_bytes": (_L, [_P]),
    # include/msclip_b200_ops.h
    "msclip_op_gemm": (_I, [_P, _L, _P, _L, _I, _I, _I, _F, _P, _P, _L, _P, _L, _I, _P]),
    "msclip_op_set_gemm_pair_mode": (None, [_I]),
    "msclip_op_layernorm": (_I, [_P, _I, _P, _P, _P, _I, _P]),
    "msclip_op_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "msclip_op_im2col_first": (_I, [_P, _I, _P, _I, _I, _I, _P]),
    "msclip_op_im2col_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P]),
    "msclip_op_conv_gemm": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P,
                                _P, _L, _I, _P]),
    "msclip_op_patch_pool": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "msclip_op_adapter_fuse_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "msclip_op_contrastive_lse": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "msclip_op_contrastive_lse_workspace": (C.c_size_t, [_I]),
    "msclip_num_keys": (_I, [_P]),
    "msclip_key_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_I), C.POINTER(_L)]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    """Every symbol the two public headers declare (used by the load/export test)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MsclipError(
                f"{LIB_PATH} is missing: build it with `python -m msclip_b200.build` "
                "(there is no Python or CPU fallback for the MS-CLIP-S path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().msclip_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise MsclipError(f"{what}: {last_error()}" if what else last_error())


def device_count() -> int:
    return int(lib().msclip_device_count())
