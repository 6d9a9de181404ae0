"""One launch of one hot op at XL shapes for a source-level ncu capture.  python tools/ncu_one.py gelu|attn|slab|pw"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402

which = sys.argv[1]
B, C, h, G, R = 1024, 768, 8, 16, 5
S = G * G + R
M = B * S
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s, sc=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g, device="cuda") * sc).to(dt)
act = rn(B, S, C)
for _ in range(2):
    if which == "gelu":
        w1, b1, hid = rn(4 * C, C, sc=0.03), rn(4 * C, dt=torch.float32), torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)
        sdp.ops.gemm(act.view(M, C), w1, hid, bias=b1, act="gelu")
    elif which == "pw":
        w = rn(C, C, sc=0.03)
        a2 = act.view(M, C)
        sdp.ops.gemm(rn(M, C), w, a2, act="gelu", residual=a2, pass_rows=(S, R))
    elif which == "attn":
        sdp.ops.attention(rn(B, S, 3 * C), torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16), h)
    elif which == "slab":
        sc = torch.empty(2 * B * G * G, device="cuda")
        lw, lb, wdw = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), rn(49, C, sc=0.1, dt=torch.float32)
        sdp.ops.ln_dwconv_slab(act, sc, lw, lb, wdw, None, torch.empty_like(act), G, G, R)
    torch.cuda.synchronize()
