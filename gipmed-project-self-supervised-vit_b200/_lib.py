"""ctypes binding of the C-ABI library ``libb200ssl.so`` (see ``include/b200ssl.h``).

There is deliberately no fallback: if the library is missing, or a call returns non-zero, the caller
gets a ``RuntimeError``.  Every entry point takes raw device pointers, sizes and a ``cudaStream_t``;
nothing here touches torch types, so the same ``.so`` can be bound from C/C++.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SSL_LIB: another build of the same library (developer A/B of kernel versions); default = the in-tree build
LIB_PATH = os.environ.get("B200SSL_LIB") or os.path.join(_HERE, "libb200ssl.so")

_lib = None

# name -> argtypes ; every function returns int (0 = ok) unless listed in _RESTYPES
_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float
SIGNATURES = {
    "b200ssl_version": [],
    "b200ssl_device_check": [],
    "b200ssl_gemm": [_P, _L, _I, _P, _L, _I, _P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P],
    "b200ssl_ln_gemm": [_P, _L, _P, _P, _F, _P, _P, _P, _P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _P],
    "b200ssl_gemm_res_ln": [_P, _L, _P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _I, _P, _P, _F, _P, _P, _P, _P],
    "b200ssl_set_gemm_cluster": [_I],
    "b200ssl_set_gemm_stationary": [_I],
    "b200ssl_set_gemm_prof": [_P],
    "b200ssl_set_attn_prof": [_P],
    "b200ssl_set_attn_stream": [_I],
    "b200ssl_set_attn_dephase": [_I],
    "b200ssl_set_gemm_wide": [_I],
    "b200ssl_set_pdl": [_I],
    "b200ssl_set_ln_bwd_staged": [_I],
    "b200ssl_layernorm_fwd": [_P, _I, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "b200ssl_layernorm_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "b200ssl_attention_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "b200ssl_attention_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "b200ssl_attention_fwd_ws": [_P, _P, _P, _I, _I, _I, _I, _F, _P, _L, _P],
    "b200ssl_attention_fwd_workspace_bytes": [_I, _I, _I],
    "b200ssl_multicrop_augment": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "b200ssl_patchify": [_P, _P, _I, _I, _I, _I, _I, _P],
    "b200ssl_assemble_tokens": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200ssl_assemble_tokens_bwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200ssl_colsum": [_P, _L, _P, _L, _I, _I, _P],
    "b200ssl_cast_f32_to_bf16": [_P, _P, _L, _P],
    "b200ssl_cast_bf16_to_f32": [_P, _P, _L, _P],
    "b200ssl_scale_rows": [_P, _P, _P, _L, _I, _P],
    "b200ssl_dropout": [_P, _P, _P, _L, _I, _F, _P, ctypes.c_uint, _P],
    "b200ssl_dropout_residual": [_P, _P, _P, _P, _L, _I, _F, _P, ctypes.c_uint, _P],
    "b200ssl_l2norm_fwd": [_P, _P, _P, _L, _I, _F, _P],
    "b200ssl_l2norm_bwd": [_P, _P, _P, _P, _L, _I, _F, _P],
    "b200ssl_weightnorm_fwd": [_P, _P, _P, _P, _L, _I, _P],
    "b200ssl_weightnorm_bwd": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "b200ssl_bn_gelu_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _I, _I, _P],
    "b200ssl_bn_gelu_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "b200ssl_zero_bytes": [_P, _L, _P],
    "b200ssl_copy_bytes": [_P, _P, _L, _P],
    "b200ssl_copy_rows": [_P, _L, _P, _L, _L, _I, _P],
    "b200ssl_add_f32": [_P, _P, _L, _P],
    "b200ssl_pos_interp": [_P, _P, _P, _I, _I, _I, _I, _P],
    "b200ssl_dino_loss_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P],
    "b200ssl_dino_loss_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P],
    "b200ssl_center_update": [_P, _P, _I, _L, _F, _P],
    "b200ssl_ema_multi_tensor": [_P, _I, _P, _P],
    "b200ssl_adamw_multi_tensor": [_P, _I, _P, _P, _F, _F, _F, _F, _P],
    "b200ssl_sumsq_multi_tensor": [_P, _I, _P, _P],
}
_RESTYPES = {"b200ssl_last_error": c_char_p, "b200ssl_attention_fwd_workspace_bytes": c_longlong}


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). b200ssl has no CPU or PyTorch fallback.")
    h = ctypes.CDLL(LIB_PATH)
    h.b200ssl_last_error.restype = c_char_p
    h.b200ssl_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = h
    return h


def last_error() -> str:
    msg = lib().b200ssl_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def exported_symbols():
    return ["b200ssl_last_error", *SIGNATURES.keys()]
