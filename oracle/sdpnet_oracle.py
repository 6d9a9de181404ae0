"""CPU oracle for the SdP-Net forward path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch-CPU tensor algebra, fp32 or fp64) of the
reference's `MainModel.forward` (`/root/reference/model.py:129-149`) and the layer
forwards it calls (`/root/reference/layers.py`).  It is NOT the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import it, and only as the checker / the timed CPU baseline.  The product path
(`sdp-net_b200/`) never imports it and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors, tests or weights, so the
oracle is pinned by executing the real reference in the build container:
`oracle/make_golden.py` imports `/root/reference/model.py`, runs it on seeded inputs and
commits the outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this
restatement against every one of those fixtures (max-abs <= 2e-6 in fp32).

The oracle works on a plain `state_dict` (name -> tensor) with the reference's key
layout, plus the reference's `model_config` dict — no nn.Module is constructed.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# model.py:28-54 -- constructor defaults of MainModel
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


def kelu(x: Tensor, a: float = 3.5) -> Tensor:
    """training_utilities.py:91-92 -- 0 below -a, x above a, smooth blend between."""
    mid = 0.5 * x * (1.0 + x / a + torch.sin(x * math.pi / a) / math.pi)
    return torch.where(x < -a, torch.zeros_like(x), torch.where(x > a, x, mid))


# model.py:13-24 -- string -> activation table ("fast_gelu" is broken upstream, omitted)
ACTIVATIONS: Dict[str, Callable[[Tensor], Tensor]] = {
    "relu": F.relu,
    "gelu": lambda t: F.gelu(t),                 # exact erf form (nn.GELU())
    "gelu_tanh": lambda t: F.gelu(t, approximate="tanh"),
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "leaky_relu": lambda t: F.leaky_relu(t, 0.01),
    "selu": F.selu,
    "none": lambda t: t,
    "kelu": kelu,
}


def _act(name_or_fn) -> Callable[[Tensor], Tensor]:
    if callable(name_or_fn):
        return name_or_fn
    return ACTIVATIONS[str(name_or_fn).lower()]


def channel_layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-6) -> Tensor:
    """layers.py:12-24 -- LayerNorm over dim 1 of NCHW, biased variance, eps inside sqrt."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    xn = (x - mu) / torch.sqrt(var + eps)
    return gamma.view(1, -1, 1, 1) * xn + beta.view(1, -1, 1, 1)


def conv_patcher(x: Tensor, w: Tensor) -> Tensor:
    """layers.py:28-42 -- Conv2d(3, C, kernel=stride=p, no bias) as an explicit patch GEMM."""
    C, cin, p, _ = w.shape
    B, _, H, W = x.shape
    gh, gw = H // p, W // p
    # unfold into [B, gh, gw, cin*p*p] with (c, dy, dx) fastest-varying order
    patches = x[:, :, : gh * p, : gw * p].reshape(B, cin, gh, p, gw, p)
    patches = patches.permute(0, 2, 4, 1, 3, 5).reshape(B, gh, gw, cin * p * p)
    out = patches @ w.reshape(C, cin * p * p).t()
    return out.permute(0, 3, 1, 2).contiguous()


def position_table(sd: Dict[str, Tensor], pre: str, H: int, W: int) -> Tensor:
    """layers.py:158-163 -- pos[c,i,j] = horizontal.weight[i,c] + vertical.weight[j,c]
    (the table *named* horizontal is indexed by the row i)."""
    eh = sd[pre + "horizontal_embedding_layer.weight"][:H]          # [H, C], rows
    ev = sd[pre + "vertical_embedding_layer.weight"][:W]            # [W, C], cols
    return eh.t()[:, :, None] + ev.t()[:, None, :]                  # [C, H, W]


def embedding_layer(x: Tensor, sd, pre: str, num_registers: int, act) -> tuple:
    """layers.py:152-168 -- add row/col tables, take register rows 0..num_registers."""
    B, C, H, W = x.shape
    x = x + position_table(sd, pre, H, W).unsqueeze(0)
    reg = sd[pre + "register_embedding_layer.weight"][: num_registers + 1]
    return _act(act)(x), reg.unsqueeze(0).expand(B, -1, -1)


def conv_embedding(x: Tensor, sd, pre: str, num_registers: int, act, ke: int) -> tuple:
    """layers.py:202-209 -- x + AvgPool_ke,stride1(bone[:, :, :H+ke-1, :W+ke-1]); activation
    covers the sum; registers are table rows 1..num_registers+1 (buffer arange(1, max+1))."""
    B, C, H, W = x.shape
    bone = sd[pre + "bone"][:, :, : H + ke - 1, : W + ke - 1]
    pos = F.avg_pool2d(bone, ke, stride=1)
    table = sd[pre + "register_embedding_layer.weight"]
    idx = torch.arange(1, num_registers + 2)
    if int(idx.max()) >= table.shape[0]:
        raise IndexError("index out of range in self")            # same failure as nn.Embedding
    return _act(act)(x + pos), table[idx].unsqueeze(0).expand(B, -1, -1)


def conv_mixer(x: Tensor, sd, pre: str, act) -> Tensor:
    """layers.py:101-104 -- LN1 -> depthwise kxk 'same' -> 1x1 C->C -> act -> +x, then
    LN2 -> 1x1 C->4C -> act -> 1x1 4C->C -> +."""
    a = _act(act)
    C = x.shape[1]
    wd = sd[pre + "conv2d.0.weight"]
    k = wd.shape[-1]
    y = channel_layernorm(x, sd[pre + "layer_norm_1.gamma"], sd[pre + "layer_norm_1.beta"])
    # padding='same' with odd k, stride 1: symmetric zero pad (k-1)/2; even k pads more on the right
    lo = (k - 1) // 2
    hi = k - 1 - lo
    y = F.pad(y, (lo, hi, lo, hi))
    y = F.conv2d(y, wd, sd.get(pre + "conv2d.0.bias"), groups=C)
    y = torch.einsum("bchw,oc->bohw", y, sd[pre + "conv2d.1.weight"][:, :, 0, 0])
    if pre + "conv2d.1.bias" in sd:
        y = y + sd[pre + "conv2d.1.bias"].view(1, -1, 1, 1)
    x1 = a(y) + x
    z = channel_layernorm(x1, sd[pre + "layer_norm_2.gamma"], sd[pre + "layer_norm_2.beta"])
    z = torch.einsum("bchw,oc->bohw", z, sd[pre + "conv1d.0.weight"][:, :, 0, 0])
    if pre + "conv1d.0.bias" in sd:
        z = z + sd[pre + "conv1d.0.bias"].view(1, -1, 1, 1)
    z = a(z)
    z = torch.einsum("bchw,oc->bohw", z, sd[pre + "conv1d.2.weight"][:, :, 0, 0])
    if pre + "conv1d.2.bias" in sd:
        z = z + sd[pre + "conv1d.2.bias"].view(1, -1, 1, 1)
    return z + x1


def _ln(x: Tensor, w: Optional[Tensor], b: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if w is not None:
        y = y * w
    if b is not None:
        y = y + b
    return y


def encoder_layer(x: Tensor, reg: Tensor, sd, pre: str, n_head: int, act) -> tuple:
    """layers.py:259-316 -- pre-LN MHSA with per-head QK LayerNorm, then pre-LN FFN, on the
    sequence [registers ; row-major patches]; eval mode (dropout / stochastic depth = id)."""
    a = _act(act)
    B, C, H, W = x.shape
    R = reg.shape[1]
    S = R + H * W
    d = C // n_head
    s = torch.cat([reg, x.flatten(2).transpose(1, 2)], dim=1)          # :271-275
    n = _ln(s, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])       # :280
    q = (n @ sd[pre + "q_proj.weight"].t()).view(B, S, n_head, d).transpose(1, 2)
    k = (n @ sd[pre + "k_proj.weight"].t()).view(B, S, n_head, d).transpose(1, 2)
    v = (n @ sd[pre + "v_proj.weight"].t()).view(B, S, n_head, d).transpose(1, 2)
    if pre + "q_norm.weight" in sd:                                    # :236-237,286
        q = _ln(q, sd[pre + "q_norm.weight"], sd[pre + "q_norm.bias"])
        k = _ln(k, sd[pre + "k_norm.weight"], sd[pre + "k_norm.bias"])
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)  # :289-298
    o = (p @ v).transpose(1, 2).reshape(B, S, C)                       # :300
    s = s + o @ sd[pre + "o_proj.weight"].t()                          # :301-303
    n2 = _ln(s, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])      # :307
    f = a(n2 @ sd[pre + "ff_linear1.weight"].t() + sd[pre + "ff_linear1.bias"])
    s = s + f @ sd[pre + "ff_linear2.weight"].t() + sd[pre + "ff_linear2.bias"]  # :308-309
    reg_out, xf = s[:, :R], s[:, R:]                                   # :311
    return xf.transpose(1, 2).reshape(B, C, H, W).contiguous(), reg_out  # :314


def classification_head(x: Tensor, reg: Tensor, sd, pre: str, from_register: bool,
                        simple: bool) -> Tensor:
    """layers.py:443-465 -- from registers: mean over R -> LN(1e-5) -> Linear [-> tanh -> Linear];
    else global average pool of x -> Linear."""
    if from_register:
        h = _ln(reg.mean(-2), sd[pre + "0.weight"], sd[pre + "0.bias"])
        h = h @ sd[pre + "1.weight"].t()
        if pre + "1.bias" in sd:
            h = h + sd[pre + "1.bias"]
        if simple:
            return h
        h = torch.tanh(h) @ sd[pre + "4.weight"].t()
        if pre + "4.bias" in sd:
            h = h + sd[pre + "4.bias"]
        return h
    h = x.mean(dim=(2, 3)) @ sd[pre + "2.weight"].t()
    if pre + "2.bias" in sd:
        h = h + sd[pre + "2.bias"]
    return h


def full_config(cfg: dict) -> dict:
    out = dict(MODEL_DEFAULTS)
    out.update(cfg)
    return out


def forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, num_registers: int = 3,
            return_raw_outputs: bool = False, dtype=torch.float32,
            stages: Optional[dict] = None):
    """model.py:129-149.  `stages`, if a dict, receives the token-major activation
    [B, R+T, C] after the embedding and after every encoder / mixer (for stage parity)."""
    cfg = full_config(cfg)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    x = x.to(dtype)
    act, eact = cfg["activation"], cfg["embedding_activation"]
    nh = cfg["n_head"]

    def note(name, xs, rs):
        if stages is not None:
            stages[name] = torch.cat([rs, xs.flatten(2).transpose(1, 2)], 1).clone()

    x = conv_patcher(x, sd["conv_init.conv.weight"])                    # model.py:134
    if cfg["conv_embedding"]:
        x, reg = conv_embedding(x, sd, "embedding_layer.", num_registers, eact,
                                cfg["conv_embedding_kernel_size"])
    else:
        x, reg = embedding_layer(x, sd, "embedding_layer.", num_registers, eact)
    note("embed", x, reg)
    for i in range(cfg["num_blocks"]):                                  # model.py:139-140
        bp = f"blocks.{i}."

        def mixers(t):
            for j in range(cfg["conv_block_num"]):
                t = conv_mixer(t, sd, f"{bp}conv_blocks.{j}.", act)
                note(f"b{i}.mixer{j}", t, reg)
            return t

        if cfg["conv_first"]:                                           # layers.py:377-386
            x = mixers(x)
            x, reg = encoder_layer(x, reg, sd, bp + "t_block.", nh, act)
            note(f"b{i}.enc", x, reg)
        else:
            x, reg = encoder_layer(x, reg, sd, bp + "t_block.", nh, act)
            note(f"b{i}.enc", x, reg)
            x = mixers(x)
    x, reg = encoder_layer(x, reg, sd, "final_block.t_block.", nh, act)  # model.py:143
    note("final", x, reg)
    logits = classification_head(x, reg, sd, "output_head.output_head.",
                                 cfg["head_output_from_register"], cfg["simple_mlp_output"])
    if return_raw_outputs:
        return logits, x, reg
    return logits


# ----------------------------------------------------------------------------------------
# Deterministic synthetic state_dicts with the reference's key layout (no reference import):
# used by GPU parity tests at sizes where no golden fixture is committed.
# ----------------------------------------------------------------------------------------
def synth_state_dict(cfg: dict, seed: int = 0, stress: bool = False) -> Dict[str, Tensor]:
    """Key layout per SURVEY.md §8(b).  stress=False mimics the reference init
    (model.py:121-126: trunc_normal std=.01 weights, N(0,1) embeddings, LN = identity);
    stress=True draws O(1/sqrt(fan_in)) weights, non-trivial LN affine and non-zero biases."""
    cfg = full_config(cfg)
    g = torch.Generator().manual_seed(seed)
    C, nh, K = cfg["embedding_dim"], cfg["n_head"], cfg["output_classes"]
    p, k, m = cfg["patch_size"], cfg["conv_kernel_size"], cfg["ff_multiplication_factor"]
    d = C // nh
    sd: Dict[str, Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def W(name, *shape, fan_in):
        if stress:
            sd[name] = rn(*shape, std=1.0 / math.sqrt(fan_in))
        else:
            sd[name] = torch.clamp(rn(*shape, std=0.01), -2.0, 2.0)

    def bias(name, n):
        sd[name] = rn(n, std=0.3 if stress else 0.02)

    def ln(wn, bn, n):
        sd[wn] = 1.0 + rn(n, std=0.25) if stress else torch.ones(n)
        sd[bn] = rn(n, std=0.25) if stress else torch.zeros(n)

    W("conv_init.conv.weight", C, 3, p, p, fan_in=3 * p * p)
    mr, (g0, g1) = cfg["max_num_registers"], cfg["max_image_size"]
    e = "embedding_layer."
    if cfg["conv_embedding"]:
        ke = cfg["conv_embedding_kernel_size"]
        sd[e + "bone"] = rn(1, C, g0 + ke, g1 + ke, std=0.5 if stress else 0.02)
        sd[e + "register"] = torch.arange(1, mr + 1, dtype=torch.int32)
        sd[e + "register_embedding_layer.weight"] = rn(mr, C)
    else:
        sd[e + "register_embeddings"] = torch.arange(mr, dtype=torch.int32)
        sd[e + "vertical_embedding"] = torch.arange(g0, dtype=torch.int32)
        sd[e + "horizontal_embedding"] = torch.arange(g1, dtype=torch.int32)
        sd[e + "register_embedding_layer.weight"] = rn(mr, C)
        sd[e + "vertical_embedding_layer.weight"] = rn(g0, C)
        sd[e + "horizontal_embedding_layer.weight"] = rn(g1, C)

    def enc(pre):
        if cfg["normalize_qv"]:
            ln(pre + "q_norm.weight", pre + "q_norm.bias", d)
            ln(pre + "k_norm.weight", pre + "k_norm.bias", d)
        for nm in ("q", "k", "v", "o"):
            W(pre + nm + "_proj.weight", C, C, fan_in=C)
        W(pre + "ff_linear1.weight", m * C, C, fan_in=C)
        bias(pre + "ff_linear1.bias", m * C)
        W(pre + "ff_linear2.weight", C, m * C, fan_in=m * C)
        bias(pre + "ff_linear2.bias", C)
        ln(pre + "norm1.weight", pre + "norm1.bias", C)
        ln(pre + "norm2.weight", pre + "norm2.bias", C)

    for i in range(cfg["num_blocks"]):
        enc(f"blocks.{i}.t_block.")
        for j in range(cfg["conv_block_num"]):
            pre = f"blocks.{i}.conv_blocks.{j}."
            W(pre + "conv2d.0.weight", C, 1, k, k, fan_in=k * k)
            if cfg["mixer_deptwise_bias"]:
                bias(pre + "conv2d.0.bias", C)
            W(pre + "conv2d.1.weight", C, C, 1, 1, fan_in=C)
            if cfg["mixer_ffn_bias"]:
                bias(pre + "conv2d.1.bias", C)
            W(pre + "conv1d.0.weight", 4 * C, C, 1, 1, fan_in=C)
            if cfg["mixer_ffn_bias"]:
                bias(pre + "conv1d.0.bias", 4 * C)
            W(pre + "conv1d.2.weight", C, 4 * C, 1, 1, fan_in=4 * C)
            if cfg["mixer_ffn_bias"]:
                bias(pre + "conv1d.2.bias", C)
            ln(pre + "layer_norm_1.gamma", pre + "layer_norm_1.beta", C)
            ln(pre + "layer_norm_2.gamma", pre + "layer_norm_2.beta", C)
    enc("final_block.t_block.")
    h = "output_head.output_head."
    hb = cfg["output_head_bias"]
    if cfg["head_output_from_register"]:
        ln(h + "0.weight", h + "0.bias", C)
        W(h + "1.weight", K, C, fan_in=C)
        if hb:
            bias(h + "1.bias", K)
        if not cfg["simple_mlp_output"]:
            W(h + "4.weight", K, K, fan_in=K)
            if hb:
                bias(h + "4.bias", K)
    else:
        W(h + "2.weight", K, C, fan_in=C)
        if hb:
            bias(h + "2.bias", K)
    return sd


def eval_metrics(logits: Tensor, labels: Tensor, label_smoothing: float = 0.0) -> dict:
    """model_test.py:70-83 -- nn.CrossEntropyLoss() (mean over rows), the reference's BCEWithLogitsLoss
    (training_utilities.py:95-107: one-hot, smooth to t*(1-ls)+ls/K, mean over all elements) and top-1."""
    x = logits.double()
    K = x.shape[1]
    lse = torch.logsumexp(x, dim=1)
    ce = (lse - x.gather(1, labels.view(-1, 1)).squeeze(1)).mean()
    t = F.one_hot(labels, K).double() * (1.0 - label_smoothing) + label_smoothing / K
    bce = (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean()
    acc = (x.argmax(1) == labels).double().mean()
    return {"cross_entropy": float(ce), "bce_with_logits": float(bce), "accuracy": float(acc)}


def flops_per_image(cfg: dict, H: int, W: int, R: int) -> float:
    """Algorithmic FLOPs (2*MAC) of one forward, SURVEY.md §8(d) formula."""
    cfg = full_config(cfg)
    C, p, k, K = cfg["embedding_dim"], cfg["patch_size"], cfg["conv_kernel_size"], cfg["output_classes"]
    m, nb, cbn = cfg["ff_multiplication_factor"], cfg["num_blocks"], cfg["conv_block_num"]
    T = (H // p) * (W // p)
    S = T + R
    patch = 2 * T * C * 3 * p * p
    mixer = 2 * T * C * k * k + 2 * T * C * C + 16 * T * C * C
    enc = 6 * S * C * C + 4 * S * S * C + 2 * S * C * C + 4 * m * S * C * C
    if cfg["head_output_from_register"] and not cfg["simple_mlp_output"]:
        head = 2 * C * K + 2 * K * K
    else:
        head = 2 * C * K
    return float(patch + nb * (cbn * mixer + enc) + enc + head)
