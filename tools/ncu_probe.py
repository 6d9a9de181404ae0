"""One launch of every hot kernel at XL shapes (B images), for ncu captures and per-kernel timing.
usage: python tools/ncu_probe.py [B] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402

ops = sdp.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
C, h, G, R, k = 768, 8, 16, 5, 7
T, S, d = G * G, G * G + R, C // h
M = B * S
dev = "cuda"
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0, dt=bf: (torch.randn(*s, generator=g, device=dev) * sc).to(dt)
act, norm, attn = rn(B, S, C), torch.empty(B, S, C, device=dev, dtype=bf), torch.empty(B, S, C, device=dev, dtype=bf)
qkv, hid = torch.empty(B, S, 3 * C, device=dev, dtype=bf), torch.empty(B, S, 4 * C, device=dev, dtype=bf)
w_qkv, w_o = rn(3 * C, C, sc=0.03), rn(C, C, sc=0.03)
w1, w2 = rn(4 * C, C, sc=0.03), rn(C, 4 * C, sc=0.02)
b1, b2 = rn(4 * C, dt=torch.float32), rn(C, dt=torch.float32)
lw, lb = rn(C, dt=torch.float32) * 0.1 + 1, rn(C, dt=torch.float32) * 0.1
qw, qb = torch.ones(d, device=dev), torch.zeros(d, device=dev)
wdw = rn(k * k, C, sc=0.1, dt=torch.float32)
a2, n2 = act.view(M, C), norm.view(M, C)


def timed(name, fn, flops=0, nbytes=0):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[1:] or ts)
    ms = ts[len(ts) // 2]                      # median of the warm repetitions
    extra = f"  (min {ts[0]:.3f})"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    if nbytes:
        extra += f"  {nbytes / ms / 1e6:8.1f} GB/s"
    print(f"{name:28s} {ms:8.3f} ms{extra}", flush=True)


timed("layernorm_rows", lambda: ops.layernorm_rows(a2, lw, lb, n2, 1e-5), nbytes=4 * M * C)
timed("ln_dwconv k7", lambda: ops.ln_dwconv(act, lw, lb, wdw, None, norm, G, G, R), flops=2 * B * T * C * k * k,
      nbytes=4 * B * T * C)
scratch = torch.empty(2 * B * T, device=dev)
timed("ln_dwconv_slab k7", lambda: ops.ln_dwconv_slab(act, scratch, lw, lb, wdw, None, norm, G, G, R), flops=2 * B * T * C * k * k,
      nbytes=4 * B * T * C)
pstats = torch.zeros(M, 6, 2, device=dev)
pstats[:, :, 0] = (act.float().view(M, 6, C // 6).sum(-1))
pstats[:, :, 1] = ((act.float().view(M, 6, C // 6) ** 2).sum(-1))
timed("ln_dwconv_slab k7 (parts in)", lambda: ops.ln_dwconv_slab(act, scratch, lw, lb, wdw, None, norm, G, G, R, producer_stats=pstats),
      flops=2 * B * T * C * k * k, nbytes=4 * B * T * C)
timed("gemm qkv+headnorm 2304x768", lambda: ops.gemm(n2, w_qkv, qkv.view(M, 3 * C), headnorm=(d, C, 1e-5, qw, qb, qw, qb)),
      flops=2 * M * 3 * C * C)
timed("gemm qkv plain 2304x768", lambda: ops.gemm(n2, w_qkv, qkv.view(M, 3 * C)), flops=2 * M * 3 * C * C)
timed("attention S261 d96", lambda: ops.attention(qkv, attn, h), flops=4 * B * h * S * S * d)
timed("gemm o +res 768x768", lambda: ops.gemm(attn.view(M, C), w_o, a2, residual=a2), flops=2 * M * C * C)
timed("gemm ff1 gelu 3072x768", lambda: ops.gemm(n2, w1, hid.view(M, 4 * C), bias=b1, act="gelu"), flops=2 * M * 4 * C * C)
timed("gemm ff1 none 3072x768", lambda: ops.gemm(n2, w1, hid.view(M, 4 * C), bias=b1), flops=2 * M * 4 * C * C)
timed("gemm ff2 +res 768x3072", lambda: ops.gemm(hid.view(M, 4 * C), w2, a2, bias=b2, residual=a2), flops=2 * M * 4 * C * C)
timed("gemm pw gelu+res mask 768", lambda: ops.gemm(n2, w_o, a2, act="gelu", residual=a2, pass_rows=(S, R)), flops=2 * M * C * C)
parts = ops.gemm_stats_parts(C, bf)
stats = torch.zeros(M, parts, 2, device=dev)
ops.row_stats(a2, stats)
sv3, tv3 = rn(3 * C, dt=torch.float32), rn(3 * C, dt=torch.float32)
sv4, tv4 = rn(4 * C, dt=torch.float32), rn(4 * C, dt=torch.float32)
timed("row_stats", lambda: ops.row_stats(a2, stats), nbytes=2 * M * C)
timed("ln_dwconv k7 (stats in)", lambda: ops.ln_dwconv(act, lw, lb, wdw, None, norm, G, G, R, stats=stats), flops=2 * B * T * C * k * k)
timed("gemm qkv fold+headnorm", lambda: ops.gemm(a2, w_qkv, qkv.view(M, 3 * C), headnorm=(d, C, 1e-5, qw, qb, qw, qb),
                                                ln_fold=(stats, 1e-5, sv3, tv3)), flops=2 * M * 3 * C * C)
timed("gemm ff1 fold gelu", lambda: ops.gemm(a2, w1, hid.view(M, 4 * C), act="gelu", ln_fold=(stats, 1e-5, sv4, tv4)),
      flops=2 * M * 4 * C * C)
timed("gemm o +res +stats", lambda: ops.gemm(attn.view(M, C), w_o, a2, residual=a2, stats_out=stats), flops=2 * M * C * C)
timed("gemm ff2 +res +stats", lambda: ops.gemm(hid.view(M, 4 * C), w2, a2, bias=b2, residual=a2, stats_out=stats),
      flops=2 * M * 4 * C * C)
timed("gemm pw gelu+res mask +stats", lambda: ops.gemm(n2, w_o, a2, act="gelu", residual=a2, pass_rows=(S, R), stats_out=stats),
      flops=2 * M * C * C)
print("done", time.strftime("%H:%M:%S"))
