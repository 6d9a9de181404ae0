"""Slab depthwise kernel in isolation at the BASELINE shapes (GPU box): ms per launch and achieved GB/s against the
algorithmic bytes (read + write of the patch rows).  python tools/dwconv_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp  # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, G, C, R) in [(1024, 16, 768, 5), (512, 14, 768, 5), (256, 14, 512, 5)]:
        T, k = G * G, 7
        act = torch.randn(B, R + T, C, device="cuda", generator=g).bfloat16()
        out = torch.empty_like(act)
        gamma, beta = torch.randn(C, device="cuda", generator=g) * 0.2 + 1, torch.randn(C, device="cuda", generator=g) * 0.2
        w = torch.randn(k * k, C, device="cuda", generator=g) * 0.1
        scratch = torch.empty(2 * B * T, device="cuda")
        parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
        pstats = torch.empty(B * (R + T), parts, 2, device="cuda")
        sdp.ops.row_stats(act, pstats)
        for name, ps in (("own stats pass", None), ("producer stats", pstats)):
            fn = lambda: sdp.ops.ln_dwconv_slab(act, scratch, gamma, beta, w, None, out, G, G, R, 1e-6, producer_stats=ps)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gb = 2 * B * T * C * 2 / 1e9
            print(f"B{B} G{G} C{C} {name}: {ms:.3f} ms/launch, {gb / ms * 1e3:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    main()
