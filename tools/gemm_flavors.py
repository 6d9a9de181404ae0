"""The five GEMM flavours of the XL step through sdp_gemm, each back to back for ~1.2 s (power-capped regime), next to
the plain kernel on the same shape: what each epilogue costs.   python tools/gemm_flavors.py [B] [seconds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp  # noqa: E402
from gemm_yardstick import sustained  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
    S, R, C, h = 261, 5, 768, 8
    M, d = B * S, C // h
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale
    x = rn(M, C)
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
    stats = torch.empty(M, parts, 2, device="cuda")
    sdp.ops.row_stats(hi, stats)
    hid = rn(M, 4 * C).bfloat16()
    qkv = torch.empty(M, 3 * C, device="cuda", dtype=torch.bfloat16)
    W = {n: (rn(N, K, scale=K ** -0.5)).bfloat16() for n, (N, K) in dict(qkv=(3 * C, C), ff1=(4 * C, C), ff2=(C, 4 * C), o=(C, C)).items()}
    s3, t3 = rn(3 * C), rn(3 * C)
    s4, t4 = rn(4 * C), rn(4 * C)
    b4, b1 = rn(4 * C), rn(C)
    hn = (d, C, 1e-5, rn(d) * 0.2 + 1, rn(d) * 0.2, rn(d) * 0.2 + 1, rn(d) * 0.2)
    hid_out = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)
    flavours = [
        ("qkv plain", 3 * C, C, lambda: sdp.ops.gemm(hi, W["qkv"], qkv)),
        ("qkv +headnorm", 3 * C, C, lambda: sdp.ops.gemm(hi, W["qkv"], qkv, headnorm=hn)),
        ("qkv +lnfold +headnorm", 3 * C, C, lambda: sdp.ops.gemm(hi, W["qkv"], qkv, headnorm=hn, ln_fold=(stats, 1e-5, s3, t3))),
        ("ff1 plain", 4 * C, C, lambda: sdp.ops.gemm(hi, W["ff1"], hid_out)),
        ("ff1 +bias +gelu", 4 * C, C, lambda: sdp.ops.gemm(hi, W["ff1"], hid_out, bias=b4, act="gelu")),
        ("ff1 +lnfold +gelu", 4 * C, C, lambda: sdp.ops.gemm(hi, W["ff1"], hid_out, act="gelu", ln_fold=(stats, 1e-5, s4, t4))),
        ("ff2 plain (separate out)", C, 4 * C, lambda: sdp.ops.gemm(hid, W["ff2"], qkv[:, :C].contiguous() if False else hid_out[:, :C])),
        ("ff2 +bias +res(hi) +stats", C, 4 * C, lambda: sdp.ops.gemm(hid, W["ff2"], hi, bias=b1, residual=hi, stats_out=stats)),
        ("ff2 +bias +res(hi+lo) +stats", C, 4 * C, lambda: sdp.ops.gemm(hid, W["ff2"], hi, bias=b1, residual=hi, stats_out=stats, residual_lo=lo, out_lo=lo)),
        ("o plain (separate out)", C, C, lambda: sdp.ops.gemm(hi, W["o"], hid_out[:, :C])),
        ("o +res(hi) +stats", C, C, lambda: sdp.ops.gemm(qkv[:, :C], W["o"], hi, residual=hi, stats_out=stats)),
        ("o +res(hi+lo) +stats", C, C, lambda: sdp.ops.gemm(qkv[:, :C], W["o"], hi, residual=hi, stats_out=stats, residual_lo=lo, out_lo=lo)),
        ("pw +gelu +res(hi+lo) +pass +stats", C, C, lambda: sdp.ops.gemm(qkv[:, :C], W["o"], hi, act="gelu", residual=hi, stats_out=stats,
                                                                       residual_lo=lo, out_lo=lo, pass_rows=(S, R))),
    ]
    only = os.environ.get("GEMM_FLAVORS_ONLY")               # e.g. "ff1,o ": the flavours whose names start with one of these
    for name, N, K, fn in flavours:
        if only and not any(name.startswith(o) for o in only.split(",")):
            continue
        # keep the stream finite over thousands of in-place accumulations
        hi.copy_(x.bfloat16())
        ms, clk = sustained(fn, secs)
        print(f"{name:36s} N{N:5d} K{K:5d}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  [{clk}]", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
