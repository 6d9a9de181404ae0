"""LayerNorm-rows bandwidth probe (GPU box): XL shape [1024*261, 768] bf16 against a device copy of the same bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdpnet_b200 as sdp
M, C = 1024 * 261, 768
x = torch.randn(M, C, device="cuda").bfloat16()
o = torch.empty_like(x)
w, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
big = torch.empty(96 << 20, device="cuda")          # L2 flush between runs: 384 MB


def t(fn, n=20):
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


ms_cp = t(lambda: o.copy_(x))
ms_ln = t(lambda: sdp.ops.layernorm_rows(x, w, b, o, 1e-5))
gb = 2 * M * C * 2 / 1e6
print(f"layernorm_rows {ms_ln:.4f} ms {gb / ms_ln:.0f} GB/s | copy_ {ms_cp:.4f} ms {gb / ms_cp:.0f} GB/s")
ref = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-5)
print("max err", float((o.float() - ref).abs().max()))
