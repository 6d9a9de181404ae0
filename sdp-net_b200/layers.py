"""Layer modules with the reference's constructor signatures, parameter names and call
conventions (reference layers.py), whose `forward` bodies run ONLY sdpnet_b200 CUDA kernels.

The modules are parameter containers + thin drivers: they own `nn.Parameter`s under exactly the
reference's names (so a reference `state_dict` loads with strict=True), and on `forward` they
pack those parameters once (bf16 or fp32), bridge the reference's NCHW / [B,R,C] tensors to the
engine's token-major layout, and launch kernels through `ops` / `engine`.  Eval only: dropout
and stochastic depth are identities (reference utility_layers.py:18,27); calling a module in
training mode raises.  `MainModel` does not call these forwards -- it runs the whole network in
token-major form through one `sdp_forward` call.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import nn

from . import ops
from .engine import Buffers, Packer, _act_name, run_encoder, run_mixer


class _KernelModule(nn.Module):
    """Common driver plumbing: precision switch, eval-only guard, lazy weight packing."""

    precision = "bf16"

    def set_precision(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        for m in self.modules():
            if isinstance(m, _KernelModule):
                m.precision = precision
                m.__dict__.pop("_pack_cache", None)
        return self

    def _guard(self, *tensors):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: sdpnet_b200 is a forward-only (eval) engine; call .eval()")
        for t in tensors:
            if t is not None and not t.is_cuda:
                raise RuntimeError("sdpnet_b200 runs on CUDA tensors only (no CPU fallback)")

    def _packed(self, build):
        key = (self.precision, tuple((p.data_ptr(), p._version) for p in self.parameters()),
               tuple((b.data_ptr(), b._version) for b in self.buffers()))
        cache = self.__dict__.get("_pack_cache")
        if cache is None or cache[0] != key:
            cache = (key, build())
            self.__dict__["_pack_cache"] = cache
        return cache[1]

    def repack(self):
        """Drop the packed-weight cache (needed after writes the version counters do not see, e.g. through `.data`)."""
        for m in self.modules():
            if isinstance(m, _KernelModule):
                m.__dict__.pop("_pack_cache", None)
        return self

    def _device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return next(self.buffers()).device


class StochasticDepth(nn.Module):
    """reference utility_layers.py:7-27; identity in eval, which is all this engine runs."""

    def __init__(self, p: float = 0.2):
        super().__init__()
        assert 0 < p < 1, "p must be a positive number or <1"
        self.p = p

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("StochasticDepth: sdpnet_b200 is forward-only (eval)")
        return x


class LayerNorm(_KernelModule):
    """Channel-first LayerNorm, reference layers.py:12-24 (eps 1e-6, biased variance)."""

    def __init__(self, embedding_dim: int, eps: float = 1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(embedding_dim))
        self.beta = nn.Parameter(torch.zeros(embedding_dim))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._guard(x)
        B, C_, H, W = x.shape
        pk = Packer(x.device, self.precision, C_)
        act = torch.empty(B, H * W, C_, dtype=pk.dtype, device=x.device)
        ops.tokens_from_nchw(x.float(), None, act)
        out = torch.empty_like(act)
        ops.layernorm_rows(act, pk._f(self.gamma), pk._f(self.beta), out, self.eps)
        y = torch.empty(B, C_, H, W, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(out, y, None, H * W, 0)
        return y


class ConvPatcher(_KernelModule):
    """reference layers.py:28-42: Conv2d(3, C, kernel=stride=patch, bias=False) as im2col + GEMM."""

    def __init__(self, embedding_dim=128, patch_size=4):
        super().__init__()
        self.conv = nn.Conv2d(in_channels=3, out_channels=embedding_dim, kernel_size=patch_size,
                              stride=patch_size, bias=False)
        self.patch_size = patch_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._guard(x)
        B, _, H, W = x.shape
        p = self.patch_size
        C_ = self.conv.out_channels
        if H % p or W % p:
            raise ValueError(f"image {H}x{W} is not divisible by patch_size {p}")
        Kp = (3 * p * p + 7) // 8 * 8

        def build():
            pk = Packer(x.device, self.precision, C_)
            return pk, pk._w(self.conv.weight, pad_k=Kp)

        pk, w = self._packed(build)
        T = (H // p) * (W // p)
        A = torch.empty(B * T, Kp, dtype=pk.dtype, device=x.device)
        ops.im2col_patches(x.contiguous(), A, p)
        out = torch.empty(B, T, C_, dtype=pk.dtype, device=x.device)
        ops.gemm(A, w, out.view(B * T, C_), K=3 * p * p)
        y = torch.empty(B, C_, H // p, W // p, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(out, y, None, T, 0)
        return y


class ConvMixer(_KernelModule):
    """reference layers.py:63-104."""

    def __init__(self, embedding_dim: int = 768, kernel_size: int = 5, activation: Callable = nn.GELU(),
                 drop_p: float = 0.0, mixer_ffn_bias: bool = True, mixer_deptwise_bias: bool = True):
        super().__init__()
        self.conv2d = nn.Sequential(
            nn.Conv2d(embedding_dim, embedding_dim, kernel_size, groups=embedding_dim, padding="same",
                      bias=mixer_deptwise_bias),
            nn.Conv2d(embedding_dim, embedding_dim, 1, bias=mixer_ffn_bias))
        act_slot = activation if isinstance(activation, nn.Module) else nn.Identity()
        self.conv1d = nn.Sequential(
            nn.Conv2d(embedding_dim, 4 * embedding_dim, 1, bias=mixer_ffn_bias),
            act_slot,
            nn.Conv2d(4 * embedding_dim, embedding_dim, 1, bias=mixer_ffn_bias))
        self.layer_norm_1 = LayerNorm(embedding_dim)
        self.layer_norm_2 = LayerNorm(embedding_dim)
        self.activation = activation
        self._act = _act_name(activation)
        self.drop_path_1 = StochasticDepth(drop_p) if drop_p > 1e-5 else nn.Identity()
        self.drop_path_2 = StochasticDepth(drop_p) if drop_p > 1e-5 else nn.Identity()
        self.embedding_dim = embedding_dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._guard(x)
        B, C_, H, W = x.shape

        def build():
            pk = Packer(x.device, self.precision, C_, act=self._act)
            return pk, pk._pack_mixer(self.state_dict(), "")

        pk, w = self._packed(build)
        bufs = Buffers(pk, B, H * W, 0)
        ops.tokens_from_nchw(x.float(), None, bufs.act)
        run_mixer(pk, w, bufs, H, W)
        y = torch.empty(B, C_, H, W, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(bufs.act, y, None, H * W, 0)
        return y


class EmbeddingLayer(_KernelModule):
    """reference layers.py:116-168.  Note the upstream naming: the table called horizontal is
    indexed by the row."""

    def __init__(self, embedding_dim: int = 768, max_num_registers: int = 5, max_image_size=[14, 14],
                 activation: Callable = None):
        super().__init__()
        self.max_num_registers = max_num_registers
        self.activation = activation if activation is not None else nn.Identity()
        self._act = _act_name(activation)
        self.register_embedding_layer = nn.Embedding(max_num_registers, embedding_dim)
        self.vertical_embedding_layer = nn.Embedding(max_image_size[0], embedding_dim)
        self.horizontal_embedding_layer = nn.Embedding(max_image_size[1], embedding_dim)
        self.register_buffer("register_embeddings", torch.arange(max_num_registers, dtype=torch.int))
        self.register_buffer("vertical_embedding", torch.arange(max_image_size[0], dtype=torch.int))
        self.register_buffer("horizontal_embedding", torch.arange(max_image_size[1], dtype=torch.int))

    def forward(self, x: torch.Tensor, num_registers: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        self._guard(x)
        B, C_, H, W = x.shape
        eh, ev = self.horizontal_embedding_layer.weight, self.vertical_embedding_layer.weight
        if H > eh.shape[0] or W > ev.shape[0] or num_registers + 1 > self.max_num_registers:
            raise IndexError("index out of range in self")
        # input-independent table assembly = parameter packing (torch indexing on the weights)
        pos = (eh[:H, None, :] + ev[None, :W, :]).detach().float().reshape(H * W, C_).contiguous()
        y = _embed_nchw(x, pos, self._act, self.precision)
        reg = self.register_embedding_layer.weight[: num_registers + 1].detach().float()
        return y, reg.unsqueeze(0).expand(B, -1, -1)


def _embed_nchw(x: torch.Tensor, pos: torch.Tensor, act_name: str, precision: str) -> torch.Tensor:
    """NCHW -> tokens, act(x + pos) in one kernel, tokens -> NCHW (fp32)."""
    B, C_, H, W = x.shape
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    act = torch.empty(B, H * W, C_, dtype=dt, device=x.device)
    ops.tokens_from_nchw(x.float(), None, act)
    ops.embed_tokens(act, pos, 0, act_name)
    y = torch.empty(B, C_, H, W, dtype=torch.float32, device=x.device)
    ops.tokens_to_nchw(act, y, None, H * W, 0)
    return y


class ConvEmbedding(_KernelModule):
    """reference layers.py:174-209."""

    def __init__(self, embedding_dim: int = 768, kernel_size: int = 5, activation: Callable = nn.GELU(),
                 max_image_size=[14, 14], max_num_registers: int = 5, seed: int = 0, trainable_bone: bool = False):
        super().__init__()
        torch.manual_seed(seed)   # upstream side effect, kept (layers.py:185)
        self.kernel_size = kernel_size
        bone = 0.02 * torch.randn(1, embedding_dim, max_image_size[0] + kernel_size, max_image_size[1] + kernel_size)
        if trainable_bone:
            self.register_parameter("bone", nn.Parameter(bone))
        else:
            self.register_buffer("bone", bone)
        self.register_buffer("register", torch.arange(1, max_num_registers + 1, dtype=torch.int))
        self.register_embedding_layer = nn.Embedding(max_num_registers, embedding_dim)
        self.activation = activation if activation is not None else nn.Identity()
        self._act = _act_name(activation)

    def forward(self, x: torch.Tensor, num_registers: int = 3):
        self._guard(x)
        B, C_, H, W = x.shape
        ke = self.kernel_size
        # input-independent: pooled bone = parameter packing
        pos = torch.nn.functional.avg_pool2d(self.bone.detach().float()[:, :, : H + ke - 1, : W + ke - 1], ke, stride=1)
        y = _embed_nchw(x, pos[0].permute(1, 2, 0).reshape(H * W, C_).contiguous(), self._act, self.precision)
        table = self.register_embedding_layer.weight
        if num_registers + 2 > table.shape[0]:
            raise IndexError("index out of range in self")
        reg = table[1: num_registers + 2].detach().float()
        return y, reg.unsqueeze(0).expand(B, -1, -1)


class EncoderLayer(_KernelModule):
    """reference layers.py:215-316."""

    def __init__(self, embedding_dim: int = 768, n_head: int = 8, activation_func: Callable = None,
                 multiplication_factor: int = 4, ff_dropout: float = 0.2, att_dropout: float = 0.2,
                 fast_att: bool = True, normalize_qv: bool = True, drop_p: float = 0.1):
        super().__init__()
        assert embedding_dim % n_head == 0, "Number of embedding_dim must be divisible by n_head"
        self.embedding_dim, self.n_head = embedding_dim, n_head
        self.head_dim = embedding_dim // n_head
        self.att_dropout, self.fast_att = att_dropout, fast_att
        self.q_norm = nn.LayerNorm(self.head_dim) if normalize_qv else nn.Identity()
        self.k_norm = nn.LayerNorm(self.head_dim) if normalize_qv else nn.Identity()
        self.drop_path1, self.drop_path2 = (StochasticDepth(drop_p), StochasticDepth(drop_p)) if drop_p > 1e-5 \
            else (nn.Identity(), nn.Identity())
        self.q_proj = nn.Linear(embedding_dim, embedding_dim, bias=False)
        self.k_proj = nn.Linear(embedding_dim, embedding_dim, bias=False)
        self.v_proj = nn.Linear(embedding_dim, embedding_dim, bias=False)
        self.o_proj = nn.Linear(embedding_dim, embedding_dim, bias=False)
        self.ff_linear1 = nn.Linear(embedding_dim, multiplication_factor * embedding_dim, bias=True)
        self.ff_linear2 = nn.Linear(multiplication_factor * embedding_dim, embedding_dim, bias=True)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.activation = activation_func if activation_func is not None else nn.functional.gelu
        self._act = _act_name(self.activation)
        self.multiplication_factor = multiplication_factor
        self.dropout = nn.Dropout(ff_dropout)

    def forward(self, x: torch.Tensor, register: torch.Tensor, mask: Optional[torch.Tensor] = None):
        self._guard(x, register)
        if mask is not None:
            raise NotImplementedError("attention masks are never passed by MainModel (model.py:140,143); unsupported")
        B, C_, H, W = x.shape
        R = register.shape[1]

        def build():
            pk = Packer(x.device, self.precision, C_, self.n_head, self._act)
            return pk, pk._pack_encoder(self.state_dict(), "")

        pk, w = self._packed(build)
        bufs = Buffers(pk, B, H * W, R, self.multiplication_factor)
        ops.tokens_from_nchw(x.float(), register.float(), bufs.act)
        run_encoder(pk, w, bufs)
        y = torch.empty(B, C_, H, W, dtype=torch.float32, device=x.device)
        r = torch.empty(B, R, C_, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(bufs.act, y, r, H * W, R)
        return y, r


class Block(_KernelModule):
    """reference layers.py:337-386."""

    def __init__(self, embedding_dim: int = 768, n_head: int = 8, conv_block_num: int = 2,
                 activation_func: Callable = nn.GELU(), multiplication_factor: int = 2, ff_dropout: float = 0.2,
                 att_dropout: float = 0.2, conv_kernel_size: int = 5, conv_activation: Callable = nn.GELU(),
                 conv_first=False, normalize_qv: bool = True, mixer_ffn_bias: bool = False,
                 mixer_deptwise_bias: bool = False, drop_p: float = 0.1, fast_att: bool = True):
        super().__init__()
        self.t_block = EncoderLayer(embedding_dim=embedding_dim, n_head=n_head, activation_func=activation_func,
                                    multiplication_factor=multiplication_factor, ff_dropout=ff_dropout,
                                    att_dropout=att_dropout, normalize_qv=normalize_qv, drop_p=drop_p,
                                    fast_att=fast_att)
        self.conv_blocks = nn.Sequential(*[
            ConvMixer(embedding_dim=embedding_dim, kernel_size=conv_kernel_size, activation=conv_activation,
                      drop_p=drop_p, mixer_deptwise_bias=mixer_deptwise_bias, mixer_ffn_bias=mixer_ffn_bias)
            for _ in range(conv_block_num)])
        self.conv_first = conv_first

    def forward(self, x: torch.Tensor, register: torch.Tensor, mask: Optional[torch.Tensor] = None):
        if not self.conv_first:
            x, register = self.t_block(x, register, mask)
            return self.conv_blocks(x), register
        x = self.conv_blocks(x)
        return self.t_block(x, register, mask)


class FinalBlock(_KernelModule):
    """reference layers.py:400-426."""

    def __init__(self, embedding_dim: int = 768, n_head: int = 8, activation_func: Callable = None,
                 multiplication_factor: int = 2, ff_dropout: float = 0.2, att_dropout: float = 0.2,
                 normalize_qv: bool = True, drop_p: float = 0.0):
        super().__init__()
        self.t_block = EncoderLayer(embedding_dim=embedding_dim, n_head=n_head, activation_func=activation_func,
                                    multiplication_factor=multiplication_factor, ff_dropout=ff_dropout,
                                    att_dropout=att_dropout, normalize_qv=normalize_qv, drop_p=drop_p)

    def forward(self, x, register, mask=None):
        return self.t_block(x, register, mask)


class ClassificationHead(_KernelModule):
    """reference layers.py:429-465."""

    def __init__(self, embedding_dim: int = 768, output_classes: int = 1000, dropout: float = 0.2,
                 from_register: bool = True, simple_output: bool = False, bias: bool = False):
        super().__init__()
        self.from_register, self.simple_output = from_register, simple_output
        self.embedding_dim, self.output_classes = embedding_dim, output_classes
        if from_register:
            if simple_output:
                self.output_head = nn.Sequential(nn.LayerNorm(embedding_dim),
                                                 nn.Linear(embedding_dim, output_classes, bias=bias))
            else:
                self.output_head = nn.Sequential(nn.LayerNorm(embedding_dim),
                                                 nn.Linear(embedding_dim, output_classes, bias=bias), nn.Tanh(),
                                                 nn.Dropout(dropout),
                                                 nn.Linear(output_classes, output_classes, bias=bias))
        else:
            self.output_head = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(),
                                             nn.Linear(embedding_dim, output_classes, bias=bias))

    def forward(self, x: torch.Tensor, registers: torch.Tensor) -> torch.Tensor:
        self._guard(x, registers)
        B, C_, H, W = x.shape
        R = registers.shape[1]
        K = self.output_classes
        Kc = (K + 7) // 8 * 8
        sd = self.state_dict()

        def build():
            pk = Packer(x.device, self.precision, C_)
            h = "output_head."
            if self.from_register:
                w = dict(ln_w=pk._f(sd[h + "0.weight"]), ln_b=pk._f(sd[h + "0.bias"]), w1=pk._w(sd[h + "1.weight"]),
                         b1=pk._f(sd.get(h + "1.bias")))
                if not self.simple_output:
                    w.update(w2=pk._w(sd[h + "4.weight"], pad_k=Kc), b2=pk._f(sd.get(h + "4.bias")))
            else:
                w = dict(w1=pk._w(sd[h + "2.weight"]), b1=pk._f(sd.get(h + "2.bias")))
            return pk, w

        pk, w = self._packed(build)
        act = torch.empty(B, R + H * W, C_, dtype=pk.dtype, device=x.device)
        ops.tokens_from_nchw(x.float(), registers.float(), act)
        pooled = torch.empty(B, C_, dtype=pk.dtype, device=x.device)
        logits = torch.empty(B, K, dtype=torch.float32, device=x.device)
        if self.from_register:
            ops.pool_ln(act, 0, R, w["ln_w"], w["ln_b"], pooled, 1e-5)
        else:
            ops.pool_ln(act, R, H * W, None, None, pooled, 0.0)
        if self.from_register and not self.simple_output:
            hbuf = torch.zeros(B, Kc, dtype=pk.dtype, device=x.device)
            ops.gemm(pooled, w["w1"], hbuf, bias=w["b1"], act="tanh", N=K)
            ops.gemm(hbuf, w["w2"], logits, bias=w["b2"], K=K)
        else:
            ops.gemm(pooled, w["w1"], logits, bias=w["b1"])
        return logits
