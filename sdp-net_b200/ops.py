"""Tensor-level wrappers over the C-ABI (include/sdpnet_b200.h).

Each function validates its torch tensors (CUDA, dtype, contiguity), passes raw device pointers
plus the current CUDA stream to the library and raises on a non-zero status.  These are the
only compute calls the package makes; they are also registered as `torch.ops.sdpnet_b200.*`.
PyTorch here is device memory + streams, nothing else.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import ACT_IDS, SDP_BF16, SDP_F32

_DT = {torch.float32: SDP_F32, torch.bfloat16: SDP_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"sdpnet_b200 supports float32 / bfloat16 tensors, got {t.dtype}") from None


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("sdpnet_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return t


def _call(name: str, ref: torch.Tensor, *args) -> None:
    """Run the C-ABI entry `name(*args, stream)` on `ref`'s device and on that device's current stream.  The library
    launches on the CUDA *current* device, so the call is made with `ref.device` current: a model on cuda:1 works
    while cuda:0 is the process default."""
    dev = ref.device
    if dev.type != "cuda":
        raise RuntimeError("sdpnet_b200 kernels need CUDA tensors (there is no CPU fallback)")
    fn = getattr(L.lib(), name)
    if torch.cuda.current_device() == dev.index:
        L.check(fn(*args, torch.cuda.current_stream(dev).cuda_stream), name)
    else:
        with torch.cuda.device(dev):
            L.check(fn(*args, torch.cuda.current_stream(dev).cuda_stream), name)


def _same_device(*tensors) -> None:
    devs = {t.device for t in tensors if t is not None}
    if len(devs) > 1:
        raise RuntimeError(f"sdpnet_b200: tensors live on different devices: {sorted(str(d) for d in devs)}")


def act_id(name) -> int:
    if isinstance(name, int):
        return name
    try:
        return ACT_IDS[str(name).lower()]
    except KeyError:
        raise ValueError(f"unknown activation {name!r}; known: {sorted(ACT_IDS)}") from None


def gemm(A: torch.Tensor, W: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, act="none", res_first: bool = False, res_mod: int = 0,
         seq_remap=(0, 0, 0), pass_rows=(0, 0), M: Optional[int] = None, N: Optional[int] = None,
         K: Optional[int] = None, headnorm=None, ln_fold=None, stats_out: Optional[torch.Tensor] = None,
         residual_lo: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epi(A[M,K] @ W[N,K]^T); see sdp_gemm in the header for the epilogue definition.  `residual_lo` / `out_lo`:
    the lo planes of a split (hi + lo) bf16 stream, same shape and pitch as `residual` / `out`."""
    _same_device(A, W, out, bias, residual, stats_out, residual_lo, out_lo)
    if A.dim() != 2 or W.dim() != 2 or out.dim() != 2:
        raise ValueError("gemm operands must be 2-D (views with a row pitch are fine)")
    if A.stride(1) != 1 or W.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("gemm operands must be K/N-contiguous")
    if A.dtype != W.dtype:
        raise TypeError("A and W must share a dtype")
    a = L.GemmArgs()
    a.A, a.lda = _p(A), A.stride(0)
    a.W, a.ldw = _p(W), W.stride(0)
    a.bias = _p(_f32(bias, "bias"))
    a.out, a.ldo = _p(out), out.stride(0)
    a.M = A.shape[0] if M is None else M
    a.N = W.shape[0] if N is None else N
    a.K = A.shape[1] if K is None else K
    if residual is not None:
        if residual.stride(-1) != 1:
            raise ValueError("residual must be contiguous in its last dim")
        a.residual, a.ldr, a.res_dtype = _p(residual), residual.stride(0), _dt(residual)
    a.dtype, a.out_dtype = _dt(A), _dt(out)
    a.act, a.res_first, a.res_mod = act_id(act), int(res_first), int(res_mod)
    a.seq_in, a.seq_out, a.seq_off = (int(v) for v in seq_remap)
    a.pass_seq, a.pass_rows = (int(v) for v in pass_rows)
    if headnorm is not None:      # (head_dim, C, eps, q_w, q_b, k_w, k_b): per-head LayerNorm on the q|k columns
        d, Cm, eps, qw, qb, kw, kb = headnorm
        a.headnorm_d, a.headnorm_C, a.headnorm_eps = int(d), int(Cm), float(eps)
        a.hn_q_w, a.hn_q_b = _p(_f32(qw, "q_norm.weight")), _p(_f32(qb, "q_norm.bias"))
        a.hn_k_w, a.hn_k_b = _p(_f32(kw, "k_norm.weight")), _p(_f32(kb, "k_norm.bias"))
    if ln_fold is not None:       # (stats [M, parts, 2], eps, s [N], t [N]): LayerNorm folded into this GEMM
        st, eps, ls, lt = ln_fold
        a.ln_stats, a.ln_parts, a.ln_eps = _p(_f32(st, "ln stats")), int(st.shape[-2]), float(eps)
        a.ln_s, a.ln_t = _p(_f32(ls, "ln_s")), _p(_f32(lt, "ln_t"))
    for lo, hi, nm in ((residual_lo, residual, "residual_lo"), (out_lo, out, "out_lo")):
        if lo is not None and (hi is None or lo.dtype != torch.bfloat16 or hi.dtype != torch.bfloat16 or lo.shape != hi.shape
                               or lo.stride() != hi.stride()):
            raise ValueError(f"{nm} must be a bfloat16 plane with the shape and strides of its hi plane")
    a.residual_lo, a.out_lo = _p(residual_lo), _p(out_lo)
    if stats_out is not None:     # [M_out, parts, 2]: the producer emits row statistics of what it stores
        a.stats_out, a.stats_parts = _p(_f32(stats_out, "stats_out")), int(stats_out.shape[-2])
    _call("sdp_gemm", A, C.byref(a))
    return out


def gemm_headnorm_ok(head_dim: int, N: int, dtype: torch.dtype) -> bool:
    return bool(L.lib().sdp_gemm_headnorm_ok(int(head_dim), int(N), _DT[dtype]))


def gemm_stats_parts(N: int, dtype: torch.dtype) -> int:
    return int(L.lib().sdp_gemm_stats_parts(int(N), _DT[dtype]))


def row_stats(x: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    """stats [M, parts, 2] fp32 <- per-row (sum, sum of squares) of x [M, C] in the producer-GEMM layout."""
    x2 = x.reshape(-1, x.shape[-1])
    _call("sdp_row_stats", x2, _p(x2), x2.stride(0), _p(_f32(stats, "stats")), int(stats.shape[-2]), x2.shape[0],
                                  x2.shape[1], _dt(x2))
    return stats


def im2col_patches(x: torch.Tensor, A: torch.Tensor, patch: int) -> torch.Tensor:
    if x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
        raise ValueError("x must be a contiguous NCHW tensor with 3 channels")
    B, _, H, W = x.shape
    _call("sdp_im2col_patches", x, _p(x), _dt(x), _p(A), _dt(A), A.stride(0), B, H, W, patch)
    return A


def fill_registers(act: torch.Tensor, table: torch.Tensor, act_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, S, Cc = act.shape
    R = table.shape[0]
    _call("sdp_fill_registers", act, _p(act), _p(act_lo), _dt(act), _p(_f32(table, "table")), B, S, R, Cc)
    return act


def layernorm_rows(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], out: torch.Tensor,
                   eps: float) -> torch.Tensor:
    x2, o2 = x.reshape(-1, x.shape[-1]), out.reshape(-1, out.shape[-1])
    if x2.dtype != o2.dtype:
        raise TypeError("layernorm_rows: in/out dtype mismatch")
    _call("sdp_layernorm_rows", x2, _p(x2), x2.stride(0), _p(_f32(w, "w")), _p(_f32(b, "b")), _p(o2),
                                       o2.stride(0), x2.shape[0], x2.shape[1], float(eps), _dt(x2))
    return out


def ln_dwconv(act: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, wdw: torch.Tensor,
              bdw: Optional[torch.Tensor], out: torch.Tensor, Gh: int, Gw: int, R: int, eps: float = 1e-6,
              stats: Optional[torch.Tensor] = None):
    B, S, Cc = act.shape
    if S != R + Gh * Gw or not act.is_contiguous() or not out.is_contiguous():
        raise ValueError("ln_dwconv: act must be contiguous [B, R + Gh*Gw, C]")
    k = int(round(wdw.shape[0] ** 0.5))      # wdw is tap-major [k*k, C]
    if wdw.dim() != 2 or k * k != wdw.shape[0] or wdw.shape[1] != Cc:
        raise ValueError("ln_dwconv: wdw must be tap-major [k*k, C]")
    _call("sdp_ln_dwconv_stats", act, _p(act), _p(_f32(stats, "stats")), 0 if stats is None else int(stats.shape[-2]),
                                        _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")), _p(_f32(wdw, "wdw")),
                                        _p(_f32(bdw, "bdw")), _p(out), B, Gh, Gw, Cc, k, R, float(eps), _dt(act))
    return out


def ln_dwconv_slab_ok(Gh: int, Gw: int, Cc: int, k: int, dtype: torch.dtype) -> bool:
    return bool(L.lib().sdp_ln_dwconv_slab_ok(Gh, Gw, Cc, k, _DT[dtype]))


def ln_dwconv_slab(act: torch.Tensor, token_stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                   wdw: torch.Tensor, bdw: Optional[torch.Tensor], out: torch.Tensor, Gh: int, Gw: int, R: int,
                   eps: float = 1e-6, producer_stats: Optional[torch.Tensor] = None):
    """Channel-stationary tensor-core variant of ln_dwconv; `token_stats` is fp32 scratch of >= 2*B*Gh*Gw elements.
    `producer_stats` ([B*S, parts, 2] fp32, the stats_out of the GEMM that produced `act`) replaces the statistics
    pass over `act`."""
    B, S, Cc = act.shape
    if S != R + Gh * Gw or not act.is_contiguous() or not out.is_contiguous():
        raise ValueError("ln_dwconv_slab: act must be contiguous [B, R + Gh*Gw, C]")
    k = int(round(wdw.shape[0] ** 0.5))
    if wdw.dim() != 2 or k * k != wdw.shape[0] or wdw.shape[1] != Cc:
        raise ValueError("ln_dwconv_slab: wdw must be tap-major [k*k, C]")
    if token_stats.dtype != torch.float32 or token_stats.numel() < 2 * B * Gh * Gw or not token_stats.is_contiguous():
        raise ValueError("ln_dwconv_slab: token_stats must be contiguous fp32 with >= 2*B*Gh*Gw elements")
    parts = 0
    if producer_stats is not None:
        if (producer_stats.dtype != torch.float32 or producer_stats.dim() != 3 or producer_stats.shape[0] != B * S
                or producer_stats.shape[2] != 2 or not producer_stats.is_contiguous()):
            raise ValueError("ln_dwconv_slab: producer_stats must be contiguous fp32 [B*S, parts, 2]")
        parts = int(producer_stats.shape[1])
    _call("sdp_ln_dwconv_slab_stats", act, _p(act), _p(producer_stats), parts, _p(token_stats), _p(_f32(gamma, "gamma")),
                                             _p(_f32(beta, "beta")), _p(_f32(wdw, "wdw")), _p(_f32(bdw, "bdw")), _p(out),
                                             B, Gh, Gw, Cc, k, R, float(eps))
    return out


def ln_dwconv_wants_stats(Gh: int, Gw: int, Cc: int, k: int, R: int, dtype: torch.dtype) -> bool:
    return bool(L.lib().sdp_ln_dwconv_wants_stats(Gh, Gw, Cc, k, R, _DT[dtype]))


def attention(qkv: torch.Tensor, out: torch.Tensor, n_head: int, qn_w=None, qn_b=None, kn_w=None, kn_b=None,
              eps: float = 1e-5, score_bound: float = 0.0) -> torch.Tensor:
    B, S, C3 = qkv.shape
    Cc = C3 // 3
    if not qkv.is_contiguous() or not out.is_contiguous() or out.shape != (B, S, Cc):
        raise ValueError("attention: qkv [B,S,3C] and out [B,S,C] must be contiguous")
    if score_bound > 0 and qn_w is None:   # q, k already normalised, |q.k|/sqrt(d) <= score_bound: one-pass softmax
        _call("sdp_attention_bounded", qkv, _p(qkv), _p(out), B, S, n_head, Cc // n_head, float(score_bound), _dt(qkv))
        return out
    _call("sdp_attention", qkv, _p(qkv), _p(_f32(qn_w, "qn_w")), _p(_f32(qn_b, "qn_b")), _p(_f32(kn_w, "kn_w")),
                                  _p(_f32(kn_b, "kn_b")), _p(out), B, S, n_head, Cc // n_head, float(eps),
                                  _dt(qkv))
    return out


def qk_score_bound(qn_w, qn_b, kn_w, kn_b) -> float:
    """Upper bound of |q . k| / sqrt(d) (nats) for q = LayerNorm(.; qn_w, qn_b), k = LayerNorm(.; kn_w, kn_b) over d
    (layers.py:286): a normalised vector has norm <= sqrt(d), so |LN(x)| <= max|w| sqrt(d) + |b|.  2 % on top for
    the bf16 rounding of q and k."""
    d = qn_w.numel()
    rd = float(d) ** 0.5
    nq = float(qn_w.detach().abs().max()) * rd + float(qn_b.detach().float().norm())
    nk = float(kn_w.detach().abs().max()) * rd + float(kn_b.detach().float().norm())
    return 1.02 * nq * nk / rd


def pool_ln(act: torch.Tensor, row0: int, nrows: int, ln_w, ln_b, out: torch.Tensor, eps: float = 1e-5,
            act_lo: Optional[torch.Tensor] = None):
    B, S, Cc = act.shape
    _call("sdp_pool_ln", act, _p(act), _p(act_lo), _dt(act), B, S, Cc, row0, nrows, _p(_f32(ln_w, "ln_w")),
                                _p(_f32(ln_b, "ln_b")), float(eps), _p(out), _dt(out), out.stride(0))
    return out


def tokens_from_nchw(x: torch.Tensor, reg: Optional[torch.Tensor], act: torch.Tensor) -> torch.Tensor:
    B, Cc, Gh, Gw = x.shape
    R = 0 if reg is None else reg.shape[1]
    x = _f32(x.contiguous(), "x")
    reg = None if reg is None else _f32(reg.contiguous(), "reg")
    _call("sdp_tokens_from_nchw", x, _p(x), _p(reg), _p(act), _dt(act), B, Cc, Gh * Gw, R)
    return act


def tokens_to_nchw(act: torch.Tensor, x: Optional[torch.Tensor], reg: Optional[torch.Tensor], T: int, R: int,
                   act_lo: Optional[torch.Tensor] = None):
    B, S, Cc = act.shape
    _call("sdp_tokens_to_nchw", act, _p(act), _p(act_lo), _dt(act), _p(_f32(x, "x")), _p(_f32(reg, "reg")), B, Cc, T, R)
    return x, reg


def embed_tokens(act: torch.Tensor, pos: torch.Tensor, R: int, act_name="none") -> torch.Tensor:
    B, S, Cc = act.shape
    _call("sdp_embed_tokens", act, _p(act), _dt(act), _p(_f32(pos, "pos")), B, S - R, R, Cc, act_id(act_name))
    return act


def eval_metrics(logits: torch.Tensor, labels: torch.Tensor, acc: torch.Tensor, label_smoothing: float = 0.0):
    """acc [5] float64 (device) += (sum CE, sum BCE-with-logits vs smoothed one-hot, #correct, #rows, #rows whose
    label is outside [0, K) -- those contribute to nothing else)."""
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
        raise TypeError("eval_metrics: logits must be float32 [B, K]")
    if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.shape[0] != logits.shape[0]:
        raise TypeError("eval_metrics: labels must be contiguous int64 [B]")
    if acc.dtype != torch.float64 or acc.numel() != 5 or not acc.is_contiguous():
        raise TypeError("eval_metrics: acc must be 5 contiguous float64 values")
    _call("sdp_eval_metrics", logits, _p(logits), logits.stride(0), _p(labels), logits.shape[0], logits.shape[1],
                                     float(label_smoothing), _p(acc))
    return acc


def val_preprocess_workspace_bytes(desc, B: int, resize, crop, general_path: bool = False) -> int:
    """Bytes of device workspace `val_preprocess` needs for these images (`desc`: _lib.ImageDesc array, host)."""
    n = int(L.lib().sdp_val_preprocess_workspace_bytes(desc, B, resize[0], resize[1], crop[0], crop[1], int(general_path)))
    if n < 0:
        raise ValueError("sdp_val_preprocess_workspace_bytes: " + L.lib().sdp_last_error().decode("utf-8", "replace"))
    return n


def val_preprocess(pixels: torch.Tensor, desc, B: int, resize, crop, mean, std, workspace: torch.Tensor,
                   out: torch.Tensor, general_path: bool = False) -> torch.Tensor:
    """Packed uint8 RGB images (device) -> out [B, 3, crop_h, crop_w] float32 / bfloat16: the reference's
    val_transforms (hf_dataset_generator.py:27-41), see sdp_val_preprocess in the header."""
    if pixels.dtype != torch.uint8 or not pixels.is_contiguous():
        raise TypeError("val_preprocess: pixels must be a contiguous uint8 tensor")
    if workspace.dtype != torch.uint8 or not workspace.is_contiguous():
        raise TypeError("val_preprocess: workspace must be a contiguous uint8 tensor")
    if tuple(out.shape) != (B, 3, crop[0], crop[1]) or not out.is_contiguous():
        raise TypeError(f"val_preprocess: out must be contiguous [B, 3, {crop[0]}, {crop[1]}]")
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    _call("sdp_val_preprocess", pixels, _p(pixels), pixels.numel(), desc, B, resize[0], resize[1], crop[0], crop[1], m, s, _p(workspace),
                                       workspace.numel(), _p(out), _dt(out), int(general_path))
    return out


def activation(x: torch.Tensor, act, force_fast: bool = False) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    _call("sdp_activation", x, _p(x), _p(y), x.numel(), act_id(act) | (0x100 if force_fast else 0), _dt(x))
    return y


def launch_count(reset: bool = False) -> int:
    return int(L.lib().sdp_launch_count(1 if reset else 0))


def device_ok() -> bool:
    return bool(L.lib().sdp_device_ok())


# ---- torch.ops registration: C++ TORCH_LIBRARY wrappers over the C-ABI (csrc/torch_ops.cpp) -------------------------
TORCH_OPS_PATH = L.LIB_PATH.replace("libsdpnet_b200.so", "libsdpnet_b200_torch.so")


def load_torch_ops() -> None:
    """Registers `torch.ops.sdpnet_b200.{gemm, layernorm_rows, ln_dwconv, attention}` (thin C++ wrappers that pass raw
    device pointers and the tensors' current CUDA stream to the C-ABI).  Raises when the wrapper library is missing."""
    import os
    if getattr(load_torch_ops, "_done", False):
        return
    if not os.path.exists(TORCH_OPS_PATH):
        raise L.SdpNetLibraryError(f"{TORCH_OPS_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L.lib()                                  # the C-ABI library first (RTLD_GLOBAL): the wrappers link against it
    torch.ops.load_library(TORCH_OPS_PATH)
    load_torch_ops._done = True


try:                                         # at import when the build is there; `load_torch_ops()` says why if it is not
    load_torch_ops()
except L.SdpNetLibraryError:  # pragma: no cover
    pass
