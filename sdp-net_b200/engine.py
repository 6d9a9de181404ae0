"""Host side of the forward engine: weight packing, workspaces and kernel sequencing.

Input is a reference-layout `state_dict` plus the reference `model_config`;
nothing of the reference's Python runs.  Two equivalent ways to run the forward:

* `Engine.forward(..)`            -- ONE `sdp_forward` C-ABI call (the sequencing lives in
                                     csrc/forward.cu); this is the product path and what
                                     bench.py times.
* `Engine.forward(.., staged=True)` -- the same kernels issued op by op from Python through
                                     `ops.*`, able to snapshot the activation after every
                                     stage (parity tests) and reused by the layer-level modules.

Activation layout everywhere: token-major `[B, S = R + T, C]`, registers first
(reference layers.py:271-275 made permanent).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn.functional as F   # weight pre-packing only (avg-pool of the conv-embedding bone)

from . import _lib as L
from . import ops

_TORCH_DT = {"bf16": torch.bfloat16, "fp32": torch.float32}

# reference model.py:28-54
MODEL_DEFAULTS = dict(
    embedding_dim=128, num_blocks=10, n_head=4, activation="gelu", conv_kernel_size=5,
    patch_size=16, ffn_dropout=0.2, attn_dropout=0.2, output_classes=1000,
    conv_block_num=2, ff_multiplication_factor=4, max_image_size=[14, 14],
    max_num_registers=5, embedding_activation="none", conv_first=True,
    head_output_from_register=False, simple_mlp_output=False, output_head_bias=False,
    normalize_qv=True, stochastic_depth_p=[0.0, 0.0], mixer_deptwise_bias=False,
    mixer_ffn_bias=False, fast_att=True, conv_embedding=False,
    conv_embedding_kernel_size=5,
)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _act_name(a) -> str:
    """Activation spec (string, nn.Module or function, as the reference ctor accepts) -> ABI name."""
    if isinstance(a, str):
        name = a.lower()
        if name == "fast_gelu":
            raise ValueError("'fast_gelu' is nn.GELU('fast') upstream, which PyTorch rejects (SURVEY.md §0.7)")
        ops.act_id(name)
        return name
    if a is None:
        return "none"
    import torch.nn as nn
    table = {nn.ReLU: "relu", nn.Tanh: "tanh", nn.Sigmoid: "sigmoid", nn.SELU: "selu", nn.Identity: "none"}
    for cls, name in table.items():
        if isinstance(a, cls):
            return name
    if isinstance(a, nn.GELU):
        return "gelu_tanh" if getattr(a, "approximate", "none") == "tanh" else "gelu"
    if isinstance(a, nn.LeakyReLU):
        if abs(a.negative_slope - 0.01) > 1e-12:
            raise ValueError("only LeakyReLU(0.01) is supported")
        return "leaky_relu"
    fname = getattr(a, "__name__", "")
    if fname in ("gelu",):
        return "gelu"
    if fname.lower() in ("kelu",):
        return "kelu"
    if fname in ("relu", "tanh", "sigmoid", "selu"):
        return fname
    raise ValueError(f"activation {a!r} has no sdpnet_b200 epilogue")


class Packer:
    """Moves parameters to the device in kernel layout: GEMM weights [N, K] in the compute dtype,
    everything else fp32.  Owns the packed tensors (`keep`)."""

    def __init__(self, device, precision: str, C_: int, n_head: int = 1, act: str = "none", ln_fold: bool = False):
        self.ln_fold = bool(ln_fold) and precision == "bf16"
        if precision not in _TORCH_DT:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        self.dtype = _TORCH_DT[precision]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sdpnet_b200 runs on CUDA devices only (no CPU fallback)")
        self.C, self.n_head, self.act = int(C_), int(n_head), act
        if precision == "bf16" and self.C % 8 != 0:
            raise ValueError("bf16 mode needs embedding_dim % 8 == 0 (16-byte TMA row pitch); use precision='fp32'")
        self.keep = []

    # -- helpers ---------------------------------------------------------------------------
    def _w(self, t: torch.Tensor, pad_k: int = 0) -> torch.Tensor:
        t = t.detach().to(self.device, torch.float32).reshape(t.shape[0], -1)
        if pad_k and t.shape[1] != pad_k:
            t = F.pad(t, (0, pad_k - t.shape[1]))
        t = t.to(self.dtype).contiguous()
        self.keep.append(t)
        return t

    def _f(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = t.detach().to(self.device, torch.float32).contiguous()
        self.keep.append(t)
        return t

    def _fold(self, W: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor]):
        """LN(x) @ W^T + b  ==  rstd * (x @ W'^T - mean * s) + t  with  W' = W diag(gamma) (bf16),
        s = rowsum(W') of the ROUNDED weights (what the tensor cores multiply), t = W beta + b."""
        W = W.detach().to(self.device, torch.float32).reshape(W.shape[0], -1)
        gamma, beta = gamma.detach().to(self.device, torch.float32), beta.detach().to(self.device, torch.float32)
        Wf = (W * gamma[None, :]).to(self.dtype).contiguous()
        s_vec = Wf.float().sum(1).contiguous()
        t_vec = (W @ beta)
        if bias is not None:
            t_vec = t_vec + bias.detach().to(self.device, torch.float32)
        t_vec = t_vec.contiguous()
        self.keep += [Wf, s_vec, t_vec]
        return Wf, s_vec, t_vec

    def _pack_encoder(self, sd, pre: str) -> dict:
        g = lambda k: sd.get(pre + k)
        w = dict(
            norm1_w=self._f(g("norm1.weight")), norm1_b=self._f(g("norm1.bias")),
            norm2_w=self._f(g("norm2.weight")), norm2_b=self._f(g("norm2.bias")),
            qn_w=self._f(g("q_norm.weight")), qn_b=self._f(g("q_norm.bias")),
            kn_w=self._f(g("k_norm.weight")), kn_b=self._f(g("k_norm.bias")),
            w_qkv=self._w(torch.cat([g("q_proj.weight"), g("k_proj.weight"), g("v_proj.weight")], 0)),
            w_o=self._w(g("o_proj.weight")),
            w_ff1=self._w(g("ff_linear1.weight")), b_ff1=self._f(g("ff_linear1.bias")),
            w_ff2=self._w(g("ff_linear2.weight")), b_ff2=self._f(g("ff_linear2.bias")),
            s_qkv=None, t_qkv=None, s_ff1=None, t_ff1=None,
            qk_score_bound=0.0,
        )
        if w["qn_w"] is not None:    # what the per-head LayerNorm lets q . k / sqrt(d) reach: attention needs no row maximum
            w["qk_score_bound"] = ops.qk_score_bound(w["qn_w"], w["qn_b"], w["kn_w"], w["kn_b"])
        if self.ln_fold:    # norm1 -> QKV and norm2 -> ff_linear1 folded (layers.py:280-284, 307-308)
            wqkv = torch.cat([g("q_proj.weight"), g("k_proj.weight"), g("v_proj.weight")], 0)
            w["w_qkv"], w["s_qkv"], w["t_qkv"] = self._fold(wqkv, g("norm1.weight"), g("norm1.bias"), None)
            w["w_ff1"], w["s_ff1"], w["t_ff1"] = self._fold(g("ff_linear1.weight"), g("norm2.weight"), g("norm2.bias"),
                                                            g("ff_linear1.bias"))
        return w

    def _pack_mixer(self, sd, pre: str) -> dict:
        g = lambda k: sd.get(pre + k)
        wd = g("conv2d.0.weight")
        w = dict(
            ln1_g=self._f(g("layer_norm_1.gamma")), ln1_b=self._f(g("layer_norm_1.beta")),
            ln2_g=self._f(g("layer_norm_2.gamma")), ln2_b=self._f(g("layer_norm_2.beta")),
            w_dw=self._f(wd.reshape(wd.shape[0], -1).t()),   # tap-major [k*k, C]
            b_dw=self._f(g("conv2d.0.bias")), conv_k=int(wd.shape[-1]),
            w_pw=self._w(g("conv2d.1.weight")), b_pw=self._f(g("conv2d.1.bias")),
            w_mlp1=self._w(g("conv1d.0.weight")), b_mlp1=self._f(g("conv1d.0.bias")),
            w_mlp2=self._w(g("conv1d.2.weight")), b_mlp2=self._f(g("conv1d.2.bias")),
            s_mlp1=None, t_mlp1=None,
        )
        if self.ln_fold:    # layer_norm_2 -> conv1d[0] folded (layers.py:103)
            w["w_mlp1"], w["s_mlp1"], w["t_mlp1"] = self._fold(g("conv1d.0.weight"), g("layer_norm_2.gamma"),
                                                               g("layer_norm_2.beta"), g("conv1d.0.bias"))
        return w


class PackedWeights(Packer):
    """The whole model's parameters, packed (see Packer)."""

    def __init__(self, cfg: dict, sd: Dict[str, torch.Tensor], device, precision: str = "bf16", ln_fold: bool = False):
        full = dict(MODEL_DEFAULTS)
        full.update(cfg)
        self.cfg = full
        C_ = int(full["embedding_dim"])
        n_head = int(full["n_head"])
        if C_ % n_head != 0:   # layers.py:229
            raise ValueError("Number of embedding_dim must be divisible by n_head")
        # folding needs whole 32-column chunks in the consumer GEMMs (N = 3C, mC, 4C) and k in {3,5,7}
        ln_fold = ln_fold and precision == "bf16" and C_ % 32 == 0 and int(full["conv_kernel_size"]) in (3, 5, 7)
        super().__init__(device, precision, C_, n_head, _act_name(full["activation"]), ln_fold)
        self.embed_act = _act_name(full["embedding_activation"])
        self._pos_cache: Dict[tuple, torch.Tensor] = {}
        self._reg_cache: Dict[int, torch.Tensor] = {}
        self.sd = sd
        self._pack(sd)

    def _pack(self, sd) -> None:
        cfg = self.cfg
        p = int(cfg["patch_size"])
        self.Kp = _round_up(3 * p * p, 8)
        self.w_patch = self._w(sd["conv_init.conv.weight"], pad_k=self.Kp)
        nb, cbn = int(cfg["num_blocks"]), int(cfg["conv_block_num"])
        self.enc = [self._pack_encoder(sd, f"blocks.{i}.t_block.") for i in range(nb)]
        self.enc.append(self._pack_encoder(sd, "final_block.t_block."))
        self.mix = [self._pack_mixer(sd, f"blocks.{i}.conv_blocks.{j}.") for i in range(nb) for j in range(cbn)]
        K = int(cfg["output_classes"])
        self.Kc = _round_up(K, 8)
        h = "output_head.output_head."
        self.head_from_register = bool(cfg["head_output_from_register"])
        self.head_simple = bool(cfg["simple_mlp_output"]) or not self.head_from_register
        self.w_head2 = self.b_head2 = self.head_ln_w = self.head_ln_b = None
        if self.head_from_register:
            self.head_ln_w, self.head_ln_b = self._f(sd[h + "0.weight"]), self._f(sd[h + "0.bias"])
            self.w_head1, self.b_head1 = self._w(sd[h + "1.weight"]), self._f(sd.get(h + "1.bias"))
            if not self.head_simple:
                self.w_head2 = self._w(sd[h + "4.weight"], pad_k=self.Kc)
                self.b_head2 = self._f(sd.get(h + "4.bias"))
        else:
            self.w_head1, self.b_head1 = self._w(sd[h + "2.weight"]), self._f(sd.get(h + "2.bias"))

    # -- input-independent tables ------------------------------------------------------------
    def pos_table(self, Gh: int, Gw: int) -> torch.Tensor:
        """[T, C] fp32.  EmbeddingLayer (layers.py:158-163): pos[i*Gw+j] = horizontal[i] + vertical[j];
        ConvEmbedding (layers.py:205): AvgPool_ke(bone[:, :, :Gh+ke-1, :Gw+ke-1])."""
        key = (Gh, Gw)
        if key not in self._pos_cache:
            sd, e = self.sd, "embedding_layer."
            if self.cfg["conv_embedding"]:
                ke = int(self.cfg["conv_embedding_kernel_size"])
                bone = sd[e + "bone"].detach().to(self.device, torch.float32)
                if Gh + ke - 1 > bone.shape[2] or Gw + ke - 1 > bone.shape[3]:
                    raise ValueError(f"grid {Gh}x{Gw} exceeds the conv-embedding bone {tuple(bone.shape[2:])}")
                pos = F.avg_pool2d(bone[:, :, : Gh + ke - 1, : Gw + ke - 1], ke, stride=1)[0]
                tab = pos.permute(1, 2, 0).reshape(Gh * Gw, self.C)
            else:
                eh = sd[e + "horizontal_embedding_layer.weight"].detach().to(self.device, torch.float32)
                ev = sd[e + "vertical_embedding_layer.weight"].detach().to(self.device, torch.float32)
                if Gh > eh.shape[0] or Gw > ev.shape[0]:
                    raise ValueError(f"grid {Gh}x{Gw} exceeds max_image_size tables ({eh.shape[0]} rows, "
                                     f"{ev.shape[0]} cols)")
                tab = (eh[:Gh, None, :] + ev[None, :Gw, :]).reshape(Gh * Gw, self.C)
            self._pos_cache[key] = tab.contiguous()
        return self._pos_cache[key]

    def reg_table(self, R: int) -> torch.Tensor:
        """[R, C] fp32: EmbeddingLayer rows 0..R-1 (layers.py:157); ConvEmbedding rows 1..R
        (layers.py:198,206) -- which is why R = max_num_registers raises IndexError there."""
        if R not in self._reg_cache:
            tab = self.sd["embedding_layer.register_embedding_layer.weight"].detach().to(self.device, torch.float32)
            lo = 1 if self.cfg["conv_embedding"] else 0
            if R < 1 or lo + R > tab.shape[0]:
                raise IndexError("index out of range in self")
            self._reg_cache[R] = tab[lo: lo + R].contiguous()
        return self._reg_cache[R]


class Buffers:
    """Workspaces for one (B, S) shape; allocated through torch, owned by the caller."""

    def __init__(self, pw: Packer, B: int, T: int, R: int, hid_mult: int = 4, Kp: int = 8, Kc: int = 8,
                 classes: int = 1, split: bool = False):
        C_, dt, dev = pw.C, pw.dtype, pw.device
        S = T + R
        hid = max(int(hid_mult), 4) * C_
        e = lambda *shape: torch.empty(*shape, dtype=dt, device=dev)
        self.B, self.T, self.R, self.S = B, T, R, S
        self.act, self.norm, self.attn = e(B, S, C_), e(B, S, C_), e(B, S, C_)
        # split residual stream (bf16): act = hi plane, act_lo = bf16(value - hi); see sdp_gemm in the header
        self.act_lo = e(B, S, C_) if (split and dt == torch.bfloat16) else None
        self.qkv = e(B, S, 3 * C_)
        self.hidden = e(B, S, hid)
        self.im2col = e(B * T, Kp)
        self.pooled = e(B, C_)
        self.head_h = torch.zeros(B, Kc, dtype=dt, device=dev)
        # row statistics (sum, sumsq): one part from sdp_row_stats for the tensor-core depthwise conv, or the
        # producer GEMMs' column parts when the LayerNorms are folded
        self.fold = bool(getattr(pw, "ln_fold", False))
        parts = max(1, ops.gemm_stats_parts(C_, dt)) if dt == torch.bfloat16 else 1
        self.stats = torch.zeros(B * S, parts, 2, dtype=torch.float32, device=dev) if dt == torch.bfloat16 else None
        # producer statistics for the tensor-core depthwise kernel: the GEMMs that feed a mixer write (sum, sumsq)
        # parts (same choice as sdp_forward); stats_fresh says whether self.stats currently describes self.act
        self.emit = False
        self.stats_fresh = False
        self.logits = torch.empty(B, classes, dtype=torch.float32, device=dev)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in vars(self).values() if isinstance(t, torch.Tensor))


# ---- op-by-op sequencing (mirrors csrc/forward.cu) ---------------------------------------------
def run_encoder(pw: Packer, w: dict, bufs: Buffers, act_name: Optional[str] = None) -> None:
    """layers.py:259-316 on bufs.act in place."""
    C_, S, B = pw.C, bufs.S, bufs.B
    M = B * S
    mult = w["w_ff1"].shape[0]
    a2, n2 = bufs.act.view(M, C_), bufs.norm.view(M, C_)
    hid = bufs.hidden.view(-1)[: M * mult].view(M, mult)
    fold = w["s_qkv"] is not None and bufs.fold
    st = bufs.stats if fold else None
    d = C_ // pw.n_head
    hn = None
    if w["qn_w"] is not None and ops.gemm_headnorm_ok(d, 3 * C_, pw.dtype):
        hn = (d, C_, 1e-5, w["qn_w"], w["qn_b"], w["kn_w"], w["kn_b"])   # q/k LayerNorm in the QKV epilogue
    xin = a2
    if not fold:
        ops.layernorm_rows(a2, w["norm1_w"], w["norm1_b"], n2, 1e-5)
        xin = n2
    ops.gemm(xin, w["w_qkv"], bufs.qkv.view(M, 3 * C_), headnorm=hn,
             ln_fold=(st, 1e-5, w["s_qkv"], w["t_qkv"]) if fold else None)
    if hn is not None:
        ops.attention(bufs.qkv, bufs.attn, pw.n_head, None, None, None, None, 1e-5, score_bound=w["qk_score_bound"])
    else:
        ops.attention(bufs.qkv, bufs.attn, pw.n_head, w["qn_w"], w["qn_b"], w["kn_w"], w["kn_b"], 1e-5)
    lo = None if bufs.act_lo is None else bufs.act_lo.view(M, C_)       # split residual stream: both planes in place
    ops.gemm(bufs.attn.view(M, C_), w["w_o"], a2, residual=a2, stats_out=st, residual_lo=lo, out_lo=lo)
    if not fold:
        ops.layernorm_rows(a2, w["norm2_w"], w["norm2_b"], n2, 1e-5)
    ops.gemm(xin, w["w_ff1"], hid, bias=None if fold else w["b_ff1"], act=act_name or pw.act,
             ln_fold=(st, 1e-5, w["s_ff1"], w["t_ff1"]) if fold else None)
    ops.gemm(hid, w["w_ff2"], a2, bias=w["b_ff2"], residual=a2, stats_out=bufs.stats if (fold or bufs.emit) else None,
             residual_lo=lo, out_lo=lo)
    bufs.stats_fresh = bool(fold or bufs.emit)


def run_mixer(pw: Packer, w: dict, bufs: Buffers, Gh: int, Gw: int, act_name: Optional[str] = None) -> None:
    """layers.py:101-104 on the patch rows of bufs.act in place (register rows pass through)."""
    C_, S, B, R = pw.C, bufs.S, bufs.B, bufs.R
    M = B * S
    act_name = act_name or pw.act
    a2, n2 = bufs.act.view(M, C_), bufs.norm.view(M, C_)
    hid = bufs.hidden.view(-1)[: M * 4 * C_].view(M, 4 * C_)
    pr = (S, R) if R > 0 else (0, 0)
    fold = w["s_mlp1"] is not None and bufs.fold
    st = bufs.stats if fold else None
    k_dw = int(round(w["w_dw"].shape[0] ** 0.5))
    have = bool((fold or bufs.emit) and bufs.stats_fresh)     # bufs.stats holds the producer GEMM's parts of act
    if bufs.stats is not None and ops.ln_dwconv_slab_ok(Gh, Gw, C_, k_dw, pw.dtype):
        # channel-stationary tensor-core kernel (same choice as sdp_forward); (mean, rstd) scratch: the head of the
        # qkv buffer (dead between two encoders) when the statistics workspace is in use, else that workspace
        scratch = bufs.qkv.view(-1).view(torch.float32) if have else bufs.stats.view(-1)
        ops.ln_dwconv_slab(bufs.act, scratch, w["ln1_g"], w["ln1_b"], w["w_dw"], w["b_dw"], bufs.norm, Gh, Gw, R, 1e-6,
                           producer_stats=bufs.stats if have else None)
    else:
        dw_stats = bufs.stats if have else None
        if not have and bufs.stats is not None and ops.ln_dwconv_wants_stats(Gh, Gw, C_, w["conv_k"], R, pw.dtype):
            ops.row_stats(a2, bufs.stats)        # token (sum, sumsq) in part 0
            dw_stats = bufs.stats
        ops.ln_dwconv(bufs.act, w["ln1_g"], w["ln1_b"], w["w_dw"], w["b_dw"], bufs.norm, Gh, Gw, R, 1e-6, stats=dw_stats)
    lo = None if bufs.act_lo is None else bufs.act_lo.view(M, C_)
    ops.gemm(n2, w["w_pw"], a2, bias=w["b_pw"], act=act_name, residual=a2, pass_rows=pr, stats_out=st,
             residual_lo=lo, out_lo=lo)
    bufs.stats_fresh = bool(fold)
    xin = a2
    if not fold:
        ops.layernorm_rows(a2, w["ln2_g"], w["ln2_b"], n2, 1e-6)
        xin = n2
    ops.gemm(xin, w["w_mlp1"], hid, bias=None if fold else w["b_mlp1"], act=act_name,
             ln_fold=(st, 1e-6, w["s_mlp1"], w["t_mlp1"]) if fold else None)
    ops.gemm(hid, w["w_mlp2"], a2, bias=w["b_mlp2"], residual=a2, pass_rows=pr,
             stats_out=bufs.stats if (fold or bufs.emit) else None, residual_lo=lo, out_lo=lo)
    bufs.stats_fresh = bool(fold or bufs.emit)


class Engine:
    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], device="cuda", precision: str = "bf16",
                 ln_fold: Optional[bool] = None, split_stream: bool = True):
        # ln_fold (bf16; default on wherever the shapes allow): the token LayerNorms in front of the QKV / FFN / mixer
        #   MLP GEMMs (layers.py:280,307 and :103) are folded into those GEMMs' epilogues; the GEMMs that write the
        #   residual stream emit the row statistics.  No stand-alone LayerNorm kernel runs.
        # split_stream (bf16; default on): the residual stream is kept as two bf16 planes hi + lo (~16 mantissa bits,
        #   the reference's autocast keeps it in fp32); False = one bf16 plane, rounded at every residual add.
        if precision not in _TORCH_DT:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        L.lib()   # fail loudly before anything else if the CUDA library is missing
        self.pw = PackedWeights(cfg, state_dict, device, precision, True if ln_fold is None else ln_fold)
        self.cfg = self.pw.cfg
        self.split_stream = bool(split_stream) and precision == "bf16"
        self._bufs: Dict[tuple, Buffers] = {}
        self._graphs: Dict[tuple, tuple] = {}          # forward_graph: (shape, dtype, registers) -> captured replay
        self._descs: Dict[tuple, tuple] = {}           # (Gh, Gw, R) -> (ModelDesc, the ctypes arrays it points into)

    # -- C-ABI model description -------------------------------------------------------------
    def _desc(self, Gh: int, Gw: int, R: int) -> L.ModelDesc:
        """The C-ABI model description for one grid / register count, built once and cached (the packed tensors it
        points into are owned by `self.pw` and never move)."""
        ent = self._descs.get((Gh, Gw, R))
        if ent is not None:
            return ent[0]
        pw, cfg = self.pw, self.cfg
        p = lambda t: None if t is None else t.data_ptr()
        enc = (L.EncoderWeights * len(pw.enc))()
        for i, w in enumerate(pw.enc):
            for name, ctype in L.EncoderWeights._fields_:
                setattr(enc[i], name, float(w[name]) if ctype is L.c_float else p(w[name]))
        mix = (L.MixerWeights * max(len(pw.mix), 1))()
        for i, w in enumerate(pw.mix):
            for name, _ in L.MixerWeights._fields_:
                setattr(mix[i], name, p(w[name]))
        d = L.ModelDesc()
        d.dtype = L.SDP_BF16 if pw.precision == "bf16" else L.SDP_F32
        d.C, d.n_head, d.num_blocks = pw.C, pw.n_head, int(cfg["num_blocks"])
        d.conv_block_num, d.ff_mult = int(cfg["conv_block_num"]), int(cfg["ff_multiplication_factor"])
        d.conv_k, d.patch, d.classes = int(cfg["conv_kernel_size"]), int(cfg["patch_size"]), int(cfg["output_classes"])
        d.act, d.embed_act = ops.act_id(pw.act), ops.act_id(pw.embed_act)
        d.conv_first = int(bool(cfg["conv_first"]))
        d.head_from_register, d.head_simple = int(pw.head_from_register), int(pw.head_simple)
        d.Kp, d.Kc = pw.Kp, pw.Kc
        d.ln_fold = int(pw.ln_fold)
        d.w_patch = p(pw.w_patch)
        d.pos_table, d.reg_table = p(pw.pos_table(Gh, Gw)), p(pw.reg_table(R))
        d.enc, d.mix = enc, mix
        d.head_ln_w, d.head_ln_b = p(pw.head_ln_w), p(pw.head_ln_b)
        d.w_head1, d.b_head1 = p(pw.w_head1), p(pw.b_head1)
        d.w_head2, d.b_head2 = p(pw.w_head2), p(pw.b_head2)
        if len(self._descs) >= 8:
            self._descs.clear()
        self._descs[(Gh, Gw, R)] = (d, enc, mix)
        return d

    def buffers(self, B: int, Gh: int, Gw: int, R: int) -> Buffers:
        key = (B, Gh * Gw, R)
        if key not in self._bufs:
            if len(self._bufs) >= 4:
                self._bufs.clear()
            self._bufs[key] = Buffers(self.pw, B, Gh * Gw, R, int(self.cfg["ff_multiplication_factor"]),
                                      self.pw.Kp, self.pw.Kc, int(self.cfg["output_classes"]), split=self.split_stream)
        return self._bufs[key]

    def _check_input(self, x: torch.Tensor, num_registers: int):
        cfg = self.cfg
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an image batch [B, 3, H, W]")
        if not x.is_cuda:
            raise RuntimeError("sdpnet_b200 runs on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("input must be float32 or bfloat16")
        p = int(cfg["patch_size"])
        B, _, H, W = x.shape
        if H % p or W % p:
            raise ValueError(f"image {H}x{W} is not divisible by patch_size {p}")
        R = int(num_registers) + 1
        if R < 1 or R > int(cfg["max_num_registers"]):
            raise IndexError("index out of range in self")   # what nn.Embedding raises upstream
        return B, H // p, W // p, R

    def forward(self, x: torch.Tensor, num_registers: int = 3, return_raw_outputs: bool = False,
                staged: bool = False, stages: Optional[dict] = None):
        """reference model.py:129-149."""
        B, Gh, Gw, R = self._check_input(x, num_registers)
        pw = self.pw
        x = x.contiguous()
        bufs = self.buffers(B, Gh, Gw, R)
        if staged or stages is not None:
            self._forward_staged(x, bufs, Gh, Gw, R, stages)
        else:
            d = self._desc(Gh, Gw, R)
            ws = L.Workspace()
            for name, _ in L.Workspace._fields_:
                t = getattr(bufs, name)
                setattr(ws, name, None if t is None else t.data_ptr())
            ops._call("sdp_forward", x, C.byref(d), C.byref(ws), x.data_ptr(), ops._dt(x), B, x.shape[2], x.shape[3], R,
                      bufs.logits.data_ptr())
        logits = bufs.logits.clone()
        if not return_raw_outputs:
            return logits
        x_raw = torch.empty(B, pw.C, Gh, Gw, dtype=torch.float32, device=x.device)
        reg = torch.empty(B, R, pw.C, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(bufs.act, x_raw, reg, Gh * Gw, R, act_lo=bufs.act_lo)
        return logits, x_raw, reg

    def forward_graph(self, x: torch.Tensor, num_registers: int = 3) -> torch.Tensor:
        """The same forward replayed from a CUDA graph (captured on first use per input shape): one graph launch instead
        of ~100 kernel launches per block.  Pays at small batches, where kernels are short next to their launch gaps
        (XL: +4.7 % at batch 256, +0.3 % at batch 1024, `tools/graph_probe.py`); bit-identical to `forward`."""
        B, Gh, Gw, R = self._check_input(x, num_registers)
        key = (tuple(x.shape), x.dtype, int(num_registers), x.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 4:
                self._graphs.clear()
            static_x = x.contiguous().clone()
            cur = torch.cuda.current_stream(x.device)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                  # lazy one-time work (attributes, tensor maps, buffers) first
                for _ in range(2):
                    self.forward(static_x, num_registers)
            cur.wait_stream(side)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(static_x, num_registers)
            ent = (graph, static_x, out, self.buffers(B, Gh, Gw, R))   # the workspaces the graph writes stay alive with it
            self._graphs[key] = ent
        graph, static_x, out, _ = ent
        static_x.copy_(x)
        graph.replay()
        return out.clone()

    def _forward_staged(self, x, bufs: Buffers, Gh: int, Gw: int, R: int, stages: Optional[dict]) -> None:
        pw, cfg = self.pw, self.cfg
        C_, T, S, B = pw.C, Gh * Gw, Gh * Gw + R, bufs.B
        p = int(cfg["patch_size"])

        def note(name):
            if stages is not None:
                stages[name] = bufs.act.float().clone() if bufs.act_lo is None else bufs.act.float() + bufs.act_lo.float()

        ops.im2col_patches(x, bufs.im2col, p)
        ops.gemm(bufs.im2col, pw.w_patch, bufs.act.view(B * S, C_), residual=pw.pos_table(Gh, Gw), res_first=True,
                 res_mod=T, act=pw.embed_act, seq_remap=(T, S, R), K=3 * p * p,
                 out_lo=None if bufs.act_lo is None else bufs.act_lo.view(B * S, C_))
        ops.fill_registers(bufs.act, pw.reg_table(R), act_lo=bufs.act_lo)
        k_dw = int(cfg["conv_kernel_size"]) if "conv_kernel_size" in cfg else 0
        sparts = bufs.stats.shape[1] if bufs.stats is not None else 0
        bufs.emit = bool(not bufs.fold and bufs.stats is not None
                         and sparts > 1 and sparts % 2 == 0 and sparts <= 16 and int(cfg["conv_block_num"]) > 0
                         and k_dw > 0 and ops.ln_dwconv_slab_ok(Gh, Gw, C_, k_dw, torch.bfloat16))
        bufs.stats_fresh = False
        if bufs.fold:
            ops.row_stats(bufs.act, bufs.stats)
            bufs.stats_fresh = True
        note("embed")
        cbn = int(cfg["conv_block_num"])
        for i in range(int(cfg["num_blocks"])):
            def mixers():
                for j in range(cbn):
                    run_mixer(pw, pw.mix[i * cbn + j], bufs, Gh, Gw)
                    note(f"b{i}.mixer{j}")
            if cfg["conv_first"]:
                mixers()
            run_encoder(pw, pw.enc[i], bufs)
            note(f"b{i}.enc")
            if not cfg["conv_first"]:
                mixers()
        run_encoder(pw, pw.enc[-1], bufs)
        note("final")
        K = int(cfg["output_classes"])
        if pw.head_from_register:
            ops.pool_ln(bufs.act, 0, R, pw.head_ln_w, pw.head_ln_b, bufs.pooled, 1e-5, act_lo=bufs.act_lo)
        else:
            ops.pool_ln(bufs.act, R, T, None, None, bufs.pooled, 0.0, act_lo=bufs.act_lo)
        if pw.head_from_register and not pw.head_simple:
            ops.gemm(bufs.pooled, pw.w_head1, bufs.head_h, bias=pw.b_head1, act="tanh", N=K)
            ops.gemm(bufs.head_h, pw.w_head2, bufs.logits, bias=pw.b_head2, K=K)
        else:
            ops.gemm(bufs.pooled, pw.w_head1, bufs.logits, bias=pw.b_head1)
