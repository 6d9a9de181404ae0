"""Run-to-run repeatability at XL shapes (GPU box): every hot op twice on identical inputs, compared bit for bit, then
the whole forward.  python tools/determinism_probe.py [B] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402

ops = sdp.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
C, h, G, R, k = 768, 8, 16, 5, 7
T, S, d = G * G, G * G + R, C // h
M = B * S
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0, dt=bf: (torch.randn(*s, generator=g, device=dev) * sc).to(dt)
act0 = rn(B, S, C)
w_qkv, w_o, w1, w2 = rn(3 * C, C, sc=0.03), rn(C, C, sc=0.03), rn(4 * C, C, sc=0.03), rn(C, 4 * C, sc=0.02)
b1, b2 = rn(4 * C, dt=torch.float32), rn(C, dt=torch.float32)
lw, lb = rn(C, dt=torch.float32) * 0.1 + 1, rn(C, dt=torch.float32) * 0.1
qw, qb = torch.ones(d, device=dev), torch.zeros(d, device=dev)
wdw = rn(k * k, C, sc=0.1, dt=torch.float32)
hid0 = rn(B, S, 4 * C, sc=0.5)
qkv0 = rn(B, S, 3 * C)
parts = ops.gemm_stats_parts(C, bf)


def check(name, fn):
    """fn() -> tuple of output tensors (fresh buffers each call)."""
    first = [t.clone() for t in fn()]
    bad = 0
    worst = 0.0
    for _ in range(reps - 1):
        again = fn()
        for a, b in zip(first, again):
            if not torch.equal(a, b):
                bad += 1
                worst = max(worst, float((a.float() - b.float()).abs().max()))
    torch.cuda.synchronize()
    print(f"{name:34s} {'REPEATABLE' if bad == 0 else f'DIFFERS in {bad} of {reps - 1} reruns, max |d| {worst:.3e}'}", flush=True)


def new(*shape, dt=bf):
    return torch.empty(*shape, device=dev, dtype=dt)


def op_ln():
    o = new(M, C); ops.layernorm_rows(act0.view(M, C), lw, lb, o, 1e-5); return (o,)
def op_slab():
    o = new(B, S, C); sc = new(2 * B * T, dt=torch.float32); ops.ln_dwconv_slab(act0, sc, lw, lb, wdw, None, o, G, G, R); return (o[:, R:],)
def op_qkv():
    o = new(M, 3 * C); ops.gemm(act0.view(M, C), w_qkv, o, headnorm=(d, C, 1e-5, qw, qb, qw, qb)); return (o,)
def op_attn():
    o = new(B, S, C); ops.attention(qkv0, o, h); return (o,)
def op_o():
    o = act0.clone().view(M, C); ops.gemm(hid0.view(M, 4 * C)[:, :C].contiguous(), w_o, o, residual=o); return (o,)
def op_ff1():
    o = new(M, 4 * C); ops.gemm(act0.view(M, C), w1, o, bias=b1, act="gelu"); return (o,)
def op_ff1_nobias():
    o = new(M, 4 * C); ops.gemm(act0.view(M, C), w1, o, act="gelu"); return (o,)
def op_ff2():
    o = act0.clone().view(M, C); st = torch.zeros(M, parts, 2, device=dev); ops.gemm(hid0.view(M, 4 * C), w2, o, bias=b2, residual=o, stats_out=st); return (o, st)
def op_pw():
    o = act0.clone().view(M, C); ops.gemm(hid0.view(M, 4 * C)[:, C:2 * C].contiguous(), w_o, o, act="gelu", residual=o, pass_rows=(S, R)); return (o,)


for name, fn in [("layernorm_rows", op_ln), ("ln_dwconv_slab", op_slab), ("gemm qkv + head-norm", op_qkv), ("attention", op_attn),
                 ("gemm o + res", op_o), ("gemm ff1 bias gelu", op_ff1), ("gemm w1 gelu", op_ff1_nobias),
                 ("gemm ff2 bias res stats", op_ff2), ("gemm pw gelu res pass", op_pw)]:
    check(name, fn)

from bench import CONFIGS, NUM_REGISTERS  # noqa: E402
cfg, _ = CONFIGS["XL"]
torch.manual_seed(0)
eng = sdp.MainModel.from_dict(**cfg).eval().to(dev).engine()
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1234)).cuda().bfloat16()
check("whole forward (sdp_forward)", lambda: (eng.forward(x, NUM_REGISTERS),))
check("whole forward (op by op)", lambda: (eng.forward(x, NUM_REGISTERS, staged=True),))

# the other BASELINE configs at their full batches, and the preprocessing kernels
for name in ("S", "M"):
    cfg_, b_ = CONFIGS[name]
    e_ = sdp.MainModel.from_dict(**cfg_).eval().to(dev).engine()
    x_ = torch.randn(b_, 3, 224, 224, generator=torch.Generator().manual_seed(99)).cuda().bfloat16()
    check(f"whole forward {name} batch {b_}", lambda: (e_.forward(x_, NUM_REGISTERS),))
gi = torch.Generator().manual_seed(5)
imgs = [torch.randint(0, 256, (*hw, 3), generator=gi, dtype=torch.uint8).numpy()
        for hw in [(375, 500), (500, 333), (480, 640), (768, 1024), (224, 224), (90, 1200)]] * 64
tf = sdp.val_transforms()
check("val_transforms 384 images", lambda: (tf(imgs),))
