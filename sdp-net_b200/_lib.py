"""ctypes binding of the C-ABI in include/sdpnet_b200.h.

The shared library is built in-tree by `build.py` (`python -m sdpnet_b200.build` or
`__graft_entry__.build()`).  There is no CPU or PyTorch fallback: if the library is missing
or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsdpnet_b200.so")

SDP_F32, SDP_BF16 = 0, 1

ACT_IDS = {
    "none": 0, "relu": 1, "gelu": 2, "gelu_tanh": 3, "tanh": 4, "sigmoid": 5,
    "leaky_relu": 6, "selu": 7, "kelu": 8,
}

c_void_p, c_int, c_i32, c_i64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float
c_float_p = C.c_void_p  # device pointers travel as integers


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_i64),
        ("W", c_void_p), ("ldw", c_i64),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_i64),
        ("out", c_void_p), ("ldo", c_i64),
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("dtype", c_i32), ("out_dtype", c_i32), ("res_dtype", c_i32),
        ("act", c_i32), ("res_first", c_i32), ("res_mod", c_i32),
        ("seq_in", c_i32), ("seq_out", c_i32), ("seq_off", c_i32),
        ("pass_seq", c_i32), ("pass_rows", c_i32),
        ("headnorm_d", c_i32), ("headnorm_C", c_i32), ("headnorm_eps", c_float),
        ("hn_q_w", c_void_p), ("hn_q_b", c_void_p), ("hn_k_w", c_void_p), ("hn_k_b", c_void_p),
        ("stats_out", c_void_p), ("stats_parts", c_i32),
        ("ln_stats", c_void_p), ("ln_parts", c_i32), ("ln_eps", c_float),
        ("ln_s", c_void_p), ("ln_t", c_void_p),
        ("residual_lo", c_void_p), ("out_lo", c_void_p),
    ]


class EncoderWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "norm1_w", "norm1_b", "norm2_w", "norm2_b", "qn_w", "qn_b", "kn_w", "kn_b",
        "w_qkv", "w_o", "w_ff1", "b_ff1", "w_ff2", "b_ff2", "s_qkv", "t_qkv", "s_ff1", "t_ff1")] + [("qk_score_bound", c_float)]


class MixerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "w_dw", "b_dw", "w_pw", "b_pw",
        "w_mlp1", "b_mlp1", "w_mlp2", "b_mlp2", "s_mlp1", "t_mlp1")]


class ModelDesc(C.Structure):
    _fields_ = [(n, c_i32) for n in (
        "dtype", "C", "n_head", "num_blocks", "conv_block_num", "ff_mult", "conv_k", "patch",
        "classes", "act", "embed_act", "conv_first", "head_from_register", "head_simple",
        "Kp", "Kc", "ln_fold")] + [
        ("w_patch", c_void_p), ("pos_table", c_void_p), ("reg_table", c_void_p),
        ("enc", C.POINTER(EncoderWeights)), ("mix", C.POINTER(MixerWeights)),
        ("head_ln_w", c_void_p), ("head_ln_b", c_void_p),
        ("w_head1", c_void_p), ("b_head1", c_void_p),
        ("w_head2", c_void_p), ("b_head2", c_void_p),
    ]


class ImageDesc(C.Structure):
    _fields_ = [("offset", c_i64), ("height", C.c_int32), ("width", C.c_int32)]


class Workspace(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "act", "act_lo", "norm", "qkv", "attn", "hidden", "im2col", "pooled", "head_h", "stats")]


# every symbol include/sdpnet_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sdp_abi_version": (c_int, []),
    "sdp_last_error": (C.c_char_p, []),
    "sdp_device_ok": (c_int, []),
    "sdp_gemm": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "sdp_gemm_headnorm_ok": (c_int, [c_int, c_int, c_int]),
    "sdp_gemm_stats_parts": (c_int, [c_int, c_int]),
    "sdp_row_stats": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_ln_dwconv_wants_stats": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "sdp_ln_dwconv_stats": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sdp_im2col_patches": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_fill_registers": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_layernorm_rows": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_float, c_int, c_void_p]),
    "sdp_ln_dwconv_slab_ok": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "sdp_ln_dwconv_slab_stats": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "sdp_ln_dwconv_slab": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "sdp_ln_dwconv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sdp_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sdp_attention_bounded": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sdp_pool_ln": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_i64, c_void_p]),
    "sdp_tokens_from_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_tokens_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_embed_tokens": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sdp_eval_metrics": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "sdp_val_preprocess_workspace_bytes": (c_i64, [C.POINTER(ImageDesc), c_int, c_int, c_int, c_int, c_int, c_int]),
    "sdp_val_preprocess": (c_int, [c_void_p, c_i64, C.POINTER(ImageDesc), c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float), C.POINTER(c_float), c_void_p, c_i64, c_void_p, c_int, c_int, c_void_p]),
    "sdp_activation": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    "sdp_forward": (c_int, [C.POINTER(ModelDesc), C.POINTER(Workspace), c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sdp_launch_count": (c_i64, [c_int]),
}

_lib = None


class SdpNetLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdpNetLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sdpnet_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)   # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.sdp_abi_version() != 7:
            raise SdpNetLibraryError("libsdpnet_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().sdp_last_error().decode("utf-8", "replace")
        raise SdpNetLibraryError(f"{what or 'sdpnet_b200 call'} failed (rc={rc}): {msg}")
