"""ctypes binding of libmsclip_b200.so (the C ABI of include/msclip_b200.h and msclip_b200_ops.h).

There is no Python / CPU fallback: if the shared library is missing this module raises, and every
compute entry point of the library fails on a box without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
# one library per MMA operand type (same sources, -DMSCLIP_FP16 for the second): bf16 is the default
_SUFFIX = os.environ.get("MSCLIP_LIB_SUFFIX", "")      # experimental build variants (msclip_b200/build.py)
LIB_PATHS = {"bf16": os.path.join(HERE, f"libmsclip_b200{_SUFFIX}.so"), "fp16": os.path.join(HERE, f"libmsclip_b200_fp16{_SUFFIX}.so")}
LIB_PATH = LIB_PATHS["bf16"]


def default_precision() -> str:
    p = os.environ.get("MSCLIP_PRECISION", "bf16").lower()
    if p not in LIB_PATHS:
        raise MsclipError(f"MSCLIP_PRECISION must be one of {sorted(LIB_PATHS)}, got {p!r}")
    return p


def torch_operand_dtype(precision: Optional[str] = None):
    import torch
    return torch.float16 if (precision or default_precision()) == "fp16" else torch.bfloat16

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
    "msclip_stage_images": (_I, [_P, _P, _I, _I, _P]),
    "msclip_encode_text": (_I, [_P, _P, _I, _P, _I, _P]),
    "msclip_set_text_trim": (_I, [_P, _I]),
    "msclip_similarity_logits": (_I, [_P, _P, _I, _P, _I, _F, _P, _P]),
    "msclip_forward": (_I, [_P, _P, _I, _P, _I, _P, _P]),
    "msclip_zeroshot_classifier": (_I, [_P, _P, _I, _I, _P, _P]),
    "msclip_zeroshot_predict": (_I, [_P, _P, _I, _P, _I, _F, _I, _P, _P, _P]),
    "msclip_comm_init": (_I, [_P, _I, _I, _I]),
    "msclip_comm_export": (_I, [_P, _P]),
    "msclip_comm_import": (_I, [_P, _P]),
    "msclip_comm_buffer": (_I, [_P, C.POINTER(_P)]),
    "msclip_comm_import_pointers": (_I, [_P, C.POINTER(_P)]),
    "msclip_contrastive_loss": (_I, [_P, _I, _F, _P, _P, _P]),
    "msclip_forward_loss": (_I, [_P, _P, _I, _P, _I, _P, _P, _P]),
    "msclip_encode_pairs": (_I, [_P, _P, _I, _P, _I, _I, _P]),
    "msclip_contrastive_loss_backward": (_I, [_P, _P, _P, _P]),
    "msclip_contrastive_loss_features": (_I, [_P, _P, _P, _I, _F, _P, _P, _P]),
    "msclip_preprocess_images": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _P, _P]),
    "msclip_tokenizer_create": (_I, [C.c_char_p, _L, C.POINTER(_P)]),
    "msclip_tokenizer_destroy": (_I, [_P]),
    "msclip_tokenizer_info": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "msclip_tokenizer_encode": (_L, [_P, C.c_char_p, _L, _P, _L]),
    "msclip_tokenizer_tokenize": (_I, [_P, C.c_char_p, _P, _I, _I, _P, _I]),
    "msclip_train_enable": (_I, [_P, _I]),
    "msclip_backward": (_I, [_P, _P, _P, _P]),
    "msclip_zero_grad": (_I, [_P, _P]),
    "msclip_taped_features": (_I, [_P, _I, _P, _I, _P]),
    "msclip_num_grads": (_I, [_P]),
    "msclip_grad_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(_L)]),
    "msclip_update_weight": (_I, [_P, C.c_char_p, _P, _P]),
    "msclip_launch_count": (_L, [_P]),
    "msclip_device_bytes": (_L, [_P]),
    # include/msclip_b200_ops.h
    "msclip_op_gemm": (_I, [_P, _L, _P, _L, _I, _I, _I, _F, _P, _P, _L, _P, _L, _I, _P]),
    "msclip_op_set_gemm_pair_mode": (None, [_I]),
    "msclip_op_gemm_ln": (_I, [_P, _L, _P, _L, _I, _I, _I, _P, _P, _L, _P, _L, _I, _I, _P, _P, _P, _L, _P, _P]),
    "msclip_op_gemm_resid_ln": (_I, [_P, _L, _P, _L, _I, _I, _I, _P, _P, _L, _P, _P, _P, _L, _P, _P]),
    "msclip_op_gemm_resid_ln_counters": (C.c_size_t, [_I]),
    "msclip_op_pack_ln_fold": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "msclip_op_layernorm": (_I, [_P, _I, _P, _P, _P, _I, _P]),
    "msclip_op_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "msclip_op_im2col_first": (_I, [_P, _I, _P, _I, _I, _I, _P]),
    "msclip_op_im2col_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P]),
    "msclip_op_conv_gemm": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P,
                                _P, _L, _I, _P]),
    "msclip_op_conv_tma": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P,
                               _P, _L, _I, _P, _P]),
    "msclip_op_conv_tma_kpad": (_I, [_I, _I, _I, _I]),
    "msclip_op_patch_pool": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "msclip_op_front_conv": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P]),
    "msclip_op_adapter_fuse_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "msclip_op_contrastive_lse": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "msclip_op_contrastive_lse_workspace": (C.c_size_t, [_I]),
    "msclip_op_wgrad": (_I, [_P, _L, _P, _L, _I, _I, _I, _P, _I, _P, _P]),
    "msclip_op_wgrad_workspace": (C.c_size_t, [_I, _I, _I]),
    "msclip_op_set_wgrad_desc": (None, [C.c_uint, C.c_uint]),
    "msclip_op_attention_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "msclip_op_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "msclip_op_qgelu_bwd": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "msclip_op_bwd_workspace": (C.c_size_t, [_I]),
    "msclip_op_adamw": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _P]),
    "msclip_num_keys": (_I, [_P]),
    "msclip_key_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_I), C.POINTER(_L)]),
}

_libs = {}


def exported_symbols():
    """Every symbol the two public headers declare (used by the load/export test)."""
    return sorted(_SIGNATURES)


def lib(precision: Optional[str] = None) -> C.CDLL:
    precision = precision or default_precision()
    if precision not in _libs:
        path = LIB_PATHS[precision]
        if not os.path.exists(path):
            raise MsclipError(
                f"{path} is missing: build it with `python -m msclip_b200.build` "
                "(there is no Python or CPU fallback for the MS-CLIP-S path)")
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _libs[precision] = handle
    return _libs[precision]


def last_error(precision: Optional[str] = None) -> str:
    return lib(precision).msclip_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "", precision: Optional[str] = None) -> None:
    if rc != 0:
        msg = last_error(precision)
        raise MsclipError(f"{what}: {msg}" if what else msg)


def device_count() -> int:
    return int(lib().msclip_device_count())
