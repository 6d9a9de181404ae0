"""Epilogue cost probe (GPU box): one GEMM shape through sdp_gemm with each epilogue ingredient switched on alone,
back to back (power-capped regime), on O(1) random data and on small 'random-init-like' data.
python tools/gemm_epi_probe.py [seconds-per-variant]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp  # noqa: E402
from gemm_yardstick import sustained  # noqa: E402


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
    M = 1024 * 261
    g = torch.Generator(device="cuda").manual_seed(0)
    for (N, K) in [(3072, 768), (768, 3072), (768, 768)]:
        for scale in (1.0, 0.02):
            A = (torch.randn(M, K, device="cuda", generator=g) * scale).bfloat16()
            W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
            bias = torch.randn(N, device="cuda", generator=g) * scale
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            res = (torch.randn(M, N, device="cuda", generator=g) * scale).bfloat16()
            fl = 2.0 * M * N * K
            variants = {
                "plain": dict(),
                "bias": dict(bias=bias),
                "gelu": dict(act="gelu"),
                "gelu+bias": dict(act="gelu", bias=bias),
                "res(sep)": dict(residual=res),
                "res(inplace)": dict(residual=out),
                "gelu+res(inplace)": dict(act="gelu", residual=out),
                "bias+res(inplace)": dict(bias=bias, residual=out),
            }
            if N == 768:
                st = torch.empty(M, sdp.ops.gemm_stats_parts(N, torch.bfloat16), 2, device="cuda")
                variants["res(inplace)+stats"] = dict(residual=out, stats_out=st)
            line = []
            for name, kw in variants.items():
                out.copy_(res)
                ms, clk = sustained(lambda: sdp.ops.gemm(A, W, out, **kw), secs)
                line.append(f"{name} {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s [{clk}]")
            ms, clk = sustained(lambda: torch.matmul(A, W.t(), out=out), secs)
            line.append(f"cuBLAS {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s [{clk}]")
            print(f"N{N} K{K} scale {scale}:\n   " + "\n   ".join(line), flush=True)
            del A, W, out, res
    return 0


if __name__ == "__main__":
    sys.exit(main())
