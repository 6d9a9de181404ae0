"""Attention kernel check + timing on the GPU box: python tools/attn_probe.py."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp  # noqa: E402


def ref(qkv, h):
    B, S, C3 = qkv.shape
    C = C3 // 3
    d = C // h
    q, k, v = [t.float().view(B, S, h, d).transpose(1, 2) for t in qkv.split(C, dim=-1)]
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1)
    return (p @ v).transpose(1, 2).reshape(B, S, C)


def bound(qkv, h):
    """max |q| max |k| / sqrt(d) over the batch: a valid (loose) bound of every score, in nats."""
    B, S, C3 = qkv.shape
    C = C3 // 3
    d = C // h
    q, k = qkv[..., :C].float().view(B, S, h, d), qkv[..., C:2 * C].float().view(B, S, h, d)
    return float(q.norm(dim=-1).max() * k.norm(dim=-1).max() / math.sqrt(d)) * 1.001


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, S, h, d) in [] if os.environ.get("ATTN_PROBE_XL_ONLY") else [(2, 128, 2, 64), (2, 16, 1, 64), (2, 5, 2, 64), (3, 201, 8, 96), (3, 261, 8, 96), (2, 256, 2, 128),
                         (2, 257, 2, 96), (2, 288, 2, 64), (2, 272, 1, 128), (5, 261, 8, 96)]:
        qkv = (torch.randn(B, S, 3 * h * d, device="cuda", generator=g) * 1.5).bfloat16()
        out = torch.full((B, S, h * d), float("nan"), device="cuda", dtype=torch.bfloat16)
        try:
            sdp.ops.attention(qkv, out, h, None, None, None, None)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"B{B} S{S} h{h} d{d}: EXC {e}")
            return 1
        r = ref(qkv, h)
        err = (out.float() - r).abs()
        out1 = torch.full_like(out, float("nan"))
        sdp.ops.attention(qkv, out1, h, score_bound=bound(qkv, h))         # one-pass softmax under a (data-derived) score bound
        torch.cuda.synchronize()
        err1 = (out1.float() - r).abs()
        print(f"B{B} S{S} h{h} d{d}: max_err {err.max().item():.4g} nan {int(torch.isnan(out.float()).sum())} "
              f"per-head {[round(err.view(B, S, h, d)[:, :, i].max().item(), 4) for i in range(min(h, 4))]}"
              f" | one-pass (bound {bound(qkv, h):.1f}): max_err {err1.max().item():.4g} nan {int(torch.isnan(out1.float()).sum())}")
    B, S, h, d = 1024, 261, 8, 96
    qkv = (torch.randn(B, S, 3 * h * d, device="cuda", generator=g)).bfloat16()
    out = torch.empty(B, S, h * d, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        sdp.ops.attention(qkv, out, h, None, None, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sdp.ops.attention(qkv, out, h, None, None, None, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * h * S * S * d
    print(f"XL attention B{B}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (useful)")
    sb = bound(qkv, h)
    for _ in range(3):
        sdp.ops.attention(qkv, out, h, score_bound=sb)
    e0.record()
    for _ in range(10):
        sdp.ops.attention(qkv, out, h, score_bound=sb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"XL attention B{B}, one-pass (bound {sb:.1f} nats): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (useful)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
