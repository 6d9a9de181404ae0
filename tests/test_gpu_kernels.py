"""Kernel-level numerics (B200 only): every C-ABI entry point against a plain PyTorch fp32
reference of the same op on the same seeded inputs.  Model-level parity against the oracle and
the golden vectors is in test_gpu_parity.py."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import sdpnet_oracle as O  # noqa: E402  (activation table for references)


@pytest.fixture(scope="module")
def sdp():
    import sdpnet_b200 as m
    m._lib.lib()
    assert m.ops.device_ok(), "needs an sm_100 device"
    return m


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    return (torch.randn(*shape, generator=_g(seed), device="cuda") * scale).to(dtype)


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


# --------------------------------------------------------------------------------------------
# GEMM: tcgen05 (bf16) and CUDA-core (fp32)
# --------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 256), (256, 512, 128), (300, 768, 768), (1000, 3072, 768),
    (1044, 768, 3072), (77, 96, 32), (64, 1000, 768), (50, 10, 32), (1024, 2304, 768),
    (130, 384, 128), (257, 128, 512), (19000, 768, 768), (8, 1000, 1000), (200, 64, 592),
]


def gemm_ref(A, W, bias=None, act="none", residual=None, res_first=False):
    v = A.float() @ W.float().t()
    if bias is not None:
        v = v + bias
    if residual is not None and res_first:
        return O.ACTIVATIONS[act](v + residual.float())
    v = O.ACTIVATIONS[act](v)
    if residual is not None:
        v = v + residual.float()
    return v


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bf16_plain(sdp, M, N, K):
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A, W, out)
    torch.cuda.synchronize()
    ref = gemm_ref(A, W)
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 1.2e-2        # bf16 output rounding (2^-8) on O(1) values
    # fp32 output isolates accumulation from output rounding
    out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    sdp.ops.gemm(A, W, out32)
    assert relerr(out32, ref) < 2e-4


@pytest.mark.parametrize("act", ["none", "gelu", "relu", "tanh", "sigmoid", "leaky_relu", "selu", "kelu", "gelu_tanh"])
def test_gemm_bf16_epilogues(sdp, act):
    M, N, K = 523, 768, 256
    A = rnd(M, K, seed=3, dtype=torch.bfloat16)
    W = rnd(N, K, seed=4, scale=2 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=5, scale=0.5)
    res = rnd(M, N, seed=6, dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    sdp.ops.gemm(A, W, out, bias=bias, act=act, residual=res)
    ref = gemm_ref(A, W, bias, act, res)
    assert (out - ref).abs().max() < 3e-3
    out_b = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A, W, out_b, bias=bias, act=act, residual=res, res_first=True)
    assert relerr(out_b, gemm_ref(A, W, bias, act, res, True)) < 1.2e-2


@pytest.mark.parametrize("M,N,K,use_bias", [(1024, 256, 64, True), (1100, 768, 768, False), (2610, 3072, 768, True),
                                            (19000, 3072, 768, True), (1025, 512, 128, True), (5000, 1024, 320, False)])
def test_gemm_bf16_gelu_lean_epilogue(sdp, M, N, K, use_bias):
    """bias + GELU, N a multiple of 256, M spanning pairs: the 16-warp lean epilogue (C -> 4C GEMMs of the model)."""
    A = rnd(M, K, seed=21, dtype=torch.bfloat16)
    W = rnd(N, K, seed=22, scale=2 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=23, scale=0.5) if use_bias else None
    out = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)     # guard rows behind M
    sdp.ops.gemm(A, W, out[:M], bias=bias, act="gelu")
    torch.cuda.synchronize()
    ref = gemm_ref(A, W, bias, "gelu")
    assert torch.isfinite(out.float()).all() and bool((out[M:] == 7.0).all())
    assert (out[:M].float() - ref).abs().max() < 1.2e-2 * max(1.0, float(ref.abs().max()))
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A, W, out2, bias=bias, act="gelu")
    assert torch.equal(out2, out[:M])                                          # run-to-run bit-exact


@pytest.mark.parametrize("B,S,R,C,K,act", [(7, 41, 5, 256, 128, "gelu"), (40, 41, 5, 256, 128, "gelu"), (33, 261, 5, 768, 768, "gelu"),
                                           (40, 41, 5, 512, 64, "none"), (9, 201, 5, 768, 768, "none"), (40, 41, 0, 256, 128, "none")])
def test_gemm_bf16_inplace_residual_and_passthrough(sdp, B, S, R, C, K, act):
    """In-place bf16 residual with pass-through rows (mixer 1x1 conv, o_proj), from one to many tiles."""
    M = B * S
    A = rnd(M, K, seed=7, dtype=torch.bfloat16)
    W = rnd(C, K, seed=8, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    x = rnd(M, C, seed=9, dtype=torch.bfloat16)
    ref = gemm_ref(A, W, None, act, x)
    rows = torch.arange(M, device="cuda") % S < R
    ref[rows] = x.float()[rows]
    out = torch.full((M + 2, C), 3.0, device="cuda", dtype=torch.bfloat16)
    out[:M] = x
    sdp.ops.gemm(A, W, out[:M], act=act, residual=out[:M], pass_rows=(S, R))
    assert relerr(out[:M], ref) < 1.2e-2
    assert torch.equal(out[:M][rows], x[rows])          # register rows bit-identical
    assert bool((out[M:] == 3.0).all())                 # nothing written behind M
    sep = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A, W, sep, act=act, residual=x)                        # residual in a separate buffer: same bits
    assert torch.equal(sep[~rows], out[:M][~rows])


def test_gemm_bf16_patch_embed_epilogue(sdp):
    """row remap [B*T] -> [B, S, C] behind R register rows + fp32 position table with row modulo."""
    B, T, R, C, K = 5, 49, 4, 128, 588
    S = T + R
    Kp = 592
    A = torch.zeros(B * T, Kp, device="cuda", dtype=torch.bfloat16)
    A[:, :K] = rnd(B * T, K, seed=10, dtype=torch.bfloat16)
    W = torch.zeros(C, Kp, device="cuda", dtype=torch.bfloat16)
    W[:, :K] = rnd(C, K, seed=11, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    pos = rnd(T, C, seed=12)
    act = torch.full((B, S, C), 7.0, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A, W, act.view(B * S, C), residual=pos, res_first=True, res_mod=T, act="gelu",
                 seq_remap=(T, S, R), K=K)
    ref = F.gelu((A[:, :K].float() @ W[:, :K].float().t()).view(B, T, C) + pos)
    assert relerr(act[:, R:], ref) < 1.2e-2
    assert (act[:, :R] == 7.0).all()                 # register rows untouched


@pytest.mark.parametrize("C,h", [(768, 8), (512, 8), (128, 4), (256, 2), (192, 2)])
def test_gemm_bf16_headnorm_epilogue(sdp, C, h):
    """QKV projection with the per-head q/k LayerNorm (layers.py:236-237,286) fused in the epilogue."""
    d, M = C // h, 333
    assert sdp.ops.gemm_headnorm_ok(d, 3 * C, torch.bfloat16)
    A = rnd(M, C, seed=60, dtype=torch.bfloat16)
    W = rnd(3 * C, C, seed=61, scale=2 / math.sqrt(C), dtype=torch.bfloat16)
    qw, qb = rnd(d, seed=62) * 0.3 + 1, rnd(d, seed=63) * 0.3
    kw, kb = rnd(d, seed=64) * 0.3 + 1, rnd(d, seed=65) * 0.3
    out = torch.empty(M, 3 * C, device="cuda", dtype=torch.float32)
    sdp.ops.gemm(A, W, out, headnorm=(d, C, 1e-5, qw, qb, kw, kb))
    raw = A.float() @ W.float().t()
    q, k, v = raw.split(C, dim=-1)
    q = F.layer_norm(q.view(M, h, d), (d,), qw, qb, 1e-5).view(M, C)
    k = F.layer_norm(k.view(M, h, d), (d,), kw, kb, 1e-5).view(M, C)
    ref = torch.cat([q, k, v], -1)
    assert (out - ref).abs().max() < 2e-3
    assert not sdp.ops.gemm_headnorm_ok(16, 96, torch.bfloat16)
    assert not sdp.ops.gemm_headnorm_ok(96, 2304, torch.float32)


@pytest.mark.parametrize("M,C,K,R", [(1305, 768, 768, 5), (603, 512, 2048, 4), (300, 128, 128, 0), (77, 32, 64, 2)])
def test_gemm_bf16_emits_row_statistics(sdp, M, C, K, R):
    """Producer side of the LayerNorm folding: per-row column-part (sum, sumsq) of the stored bf16 values;
    pass-through (register) rows keep their previous statistics."""
    S = 9 if R else 0
    A = rnd(M, K, seed=70, dtype=torch.bfloat16)
    W = rnd(C, K, seed=71, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    x = (rnd(M, C, seed=72) * 2 + 0.7).to(torch.bfloat16)
    parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
    assert parts >= 2 and parts % 2 == 0
    stats = torch.full((M, parts, 2), -7.0, device="cuda")
    out = x.clone()
    sdp.ops.gemm(A, W, out, residual=out, act="gelu", pass_rows=(S, R) if R else (0, 0), stats_out=stats)
    live = torch.ones(M, dtype=torch.bool, device="cuda") if not R else (torch.arange(M, device="cuda") % S >= R)
    got = stats.sum(1)
    o = out.float()
    assert (got[live, 0] - o[live].sum(1)).abs().max() < 2e-3 * (1 + o[live].abs().sum(1).max())
    assert (got[live, 1] - (o[live] ** 2).sum(1)).abs().max() < 2e-3 * (1 + (o[live] ** 2).sum(1).max())
    assert (stats[~live] == -7.0).all()
    # stand-alone producer agrees
    st2 = torch.empty_like(stats)
    sdp.ops.row_stats(out, st2)
    assert (st2.sum(1)[live] - got[live]).abs().max() < 2e-3 * (1 + got[live].abs().max())
    assert (st2[:, 1:] == 0).all()


@pytest.mark.parametrize("M", [777, 2610])          # below / above the CTA-pair threshold (the lean GELU epilogue needs pairs)
@pytest.mark.parametrize("C,N,act,hn", [(768, 3072, "gelu", False), (768, 2304, "none", True), (512, 1536, "none", True),
                                        (128, 512, "relu", False), (32, 96, "none", False), (512, 2048, "gelu", False)])
def test_gemm_bf16_layernorm_fold(sdp, C, N, act, hn, M):
    """Consumer side: LN(x) @ W^T + b computed as rstd * (x @ W'^T - mean * s) + t from the row statistics."""
    eps = 1e-5
    x = (rnd(M, C, seed=73) * 1.7 + 0.9).to(torch.bfloat16)
    W = rnd(N, C, seed=74, scale=1 / math.sqrt(C))
    gamma, beta, bias = rnd(C, seed=75) * 0.3 + 1, rnd(C, seed=76) * 0.3, rnd(N, seed=77) * 0.2
    Wf = (W * gamma[None, :]).to(torch.bfloat16)
    s_vec, t_vec = Wf.float().sum(1).contiguous(), (W @ beta + bias).contiguous()
    parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
    stats = torch.empty(M, parts, 2, device="cuda")
    sdp.ops.row_stats(x, stats)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    kw = {}
    h = 0
    if hn:
        h = 8
        d = (N // 3) // h
        qw, qb, kw_, kb = rnd(d, seed=78) * 0.3 + 1, rnd(d, seed=79) * 0.3, rnd(d, seed=80) * 0.3 + 1, rnd(d, seed=81) * 0.3
        kw["headnorm"] = (d, N // 3, 1e-5, qw, qb, kw_, kb)
    sdp.ops.gemm(x, Wf, out, act=act, ln_fold=(stats, eps, s_vec, t_vec), **kw)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, eps) @ W.t() + bias
    if hn:
        Cq = N // 3
        q, k, v = ref.split(Cq, dim=-1)
        q = F.layer_norm(q.view(M, h, d), (d,), qw, qb, 1e-5).view(M, Cq)
        k = F.layer_norm(k.view(M, h, d), (d,), kw_, kb, 1e-5).view(M, Cq)
        ref = torch.cat([q, k, v], -1)
    ref = O.ACTIVATIONS[act](ref)
    assert relerr(out, ref) < 1.5e-2



def test_gemm_bf16_layernorm_fold_large_mean_rows(sdp):
    """Rows whose channel mean is far from zero (|mean| = 8 std): the one-pass fp32 (sum, sumsq) statistics of the bf16
    values and the `acc - mean * s` cancellation in the consumer epilogue stay at bf16 level.  (The producer sums
    bf16 values in fp32: mean^2 / var can reach 2^16 before the data itself stops resolving the variance, which leaves
    the fp32 one-pass form ~1e-3 relative in the variance -- the reason no shifted sums are emitted.)"""
    M, C, N, eps = 2610, 768, 3072, 1e-5
    base = rnd(M, 1, seed=83) * 3.0 + 8.0
    x = (rnd(M, C, seed=84) + base).to(torch.bfloat16)                    # per-row mean ~ 8 +- 3, std 1
    W = rnd(N, C, seed=85, scale=1 / math.sqrt(C))
    gamma, beta, bias = rnd(C, seed=86) * 0.3 + 1, rnd(C, seed=87) * 0.3, rnd(N, seed=88) * 0.2
    Wf = (W * gamma[None, :]).to(torch.bfloat16)
    s_vec, t_vec = Wf.float().sum(1).contiguous(), (W @ beta + bias).contiguous()
    # statistics as the producer GEMM emits them: in-place residual epilogue over the split stream
    parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
    stats = torch.empty(M, parts, 2, device="cuda")
    hi, lo = _split(x.float())
    A0 = torch.zeros(M, 64, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(A0, torch.zeros(C, 64, device="cuda", dtype=torch.bfloat16), hi, residual=hi, residual_lo=lo, out_lo=lo, stats_out=stats)
    assert torch.equal(hi, x)                                              # adding zero changes nothing
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sdp.ops.gemm(hi, Wf, out, act="gelu", ln_fold=(stats, eps, s_vec, t_vec))
    ref = torch.nn.functional.gelu(F.layer_norm(x.float(), (C,), gamma, beta, eps) @ W.t() + bias)
    assert relerr(out, ref) < 2e-2
    mean = stats[:, :, 0].sum(1) / C
    var = stats[:, :, 1].sum(1) / C - mean * mean
    assert float(((var - x.float().var(1, unbiased=False)).abs() / x.float().var(1, unbiased=False)).max()) < 2e-3


def _split(x):
    """fp32 -> (hi, lo) bf16 planes of the split residual stream."""
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


@pytest.mark.parametrize("B,S,R,C,K,act,bias", [
    (7, 41, 5, 256, 128, "gelu", False),      # one CTA, 2-slot ring
    (40, 41, 5, 256, 128, "gelu", True),      # CTA pairs, 3-slot ring
    (33, 261, 5, 768, 768, "gelu", False),    # mixer 1x1 conv at XL width
    (9, 201, 5, 768, 3072, "none", True),     # ff2 / mlp2
    (300, 261, 5, 768, 768, "none", False),   # many tiles per CTA: the prefetch cursor crosses tile boundaries
    (40, 41, 0, 512, 64, "none", False),      # S-size width, no pass-through rows
    (25, 50, 3, 128, 256, "relu", False),     # BN = 128
    (25, 50, 3, 96, 64, "none", False),       # N not a multiple of the column half: ragged last chunk range
    (9, 21, 2, 40, 64, "tanh", True),         # N % 32 != 0: direct-store fallback, same contract
])
def test_gemm_bf16_split_stream_inplace(sdp, B, S, R, C, K, act, bias):
    """The split (hi + lo) residual stream through the in-place residual epilogue: out_hi + out_lo carries the fp32
    sum to ~16 bits, pass-through rows keep both planes bit for bit, statistics describe the hi plane."""
    M = B * S
    A = rnd(M, K, seed=7, dtype=torch.bfloat16)
    W = rnd(C, K, seed=8, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bvec = rnd(C, seed=10) * 0.3 if bias else None
    x = rnd(M, C, seed=9) * 1.5 + 0.3
    hi, lo = _split(x)
    x16 = hi.float() + lo.float()
    ref = gemm_ref(A, W, bvec, act, x16)
    rows = torch.arange(M, device="cuda") % S < R if R else torch.zeros(M, dtype=torch.bool, device="cuda")
    oh = torch.full((M + 3, C), 3.0, device="cuda", dtype=torch.bfloat16)
    ol = torch.full((M + 3, C), 5.0, device="cuda", dtype=torch.bfloat16)
    oh[:M], ol[:M] = hi, lo
    parts = sdp.ops.gemm_stats_parts(C, torch.bfloat16)
    stats = torch.full((M, parts, 2), -7.0, device="cuda") if C % 32 == 0 else None
    sdp.ops.gemm(A, W, oh[:M], bias=bvec, act=act, residual=oh[:M], pass_rows=(S, R) if R else (0, 0),
                 residual_lo=ol[:M], out_lo=ol[:M], stats_out=stats)
    got = oh[:M].float() + ol[:M].float()
    live = ~rows
    scale = float(ref.abs().max())
    assert float((got[live] - ref[live]).abs().max()) < 1e-4 * scale          # bf16 alone would be ~4e-3 * scale
    assert torch.equal(oh[:M][live], ref[live].to(torch.bfloat16)) or \
        float((oh[:M][live].float() - ref[live]).abs().max()) < 8e-3 * scale      # hi = bf16(v) (ties may differ by an ulp)
    assert float((ol[:M][live].float()).abs().max()) <= 2.0 ** -8 * scale       # lo is a rounding remainder
    assert torch.equal(oh[:M][rows], hi[rows]) and torch.equal(ol[:M][rows], lo[rows])
    assert bool((oh[M:] == 3.0).all()) and bool((ol[M:] == 5.0).all())            # nothing written behind M
    if stats is not None:
        h = oh[:M].float()
        g = stats.sum(1)
        assert (g[live, 0] - h[live].sum(1)).abs().max() < 2e-3 * (1 + h[live].abs().sum(1).max())
        assert (g[live, 1] - (h[live] ** 2).sum(1)).abs().max() < 2e-3 * (1 + (h[live] ** 2).sum(1).max())
        assert (stats[rows] == -7.0).all()
    # hi-only in-place (no lo plane) through the same epilogue equals the separate-residual result bit for bit
    o2 = hi.clone()
    sdp.ops.gemm(A, W, o2, bias=bvec, act=act, residual=o2, pass_rows=(S, R) if R else (0, 0))
    sep = torch.empty_like(hi)
    sdp.ops.gemm(A, W, sep, bias=bvec, act=act, residual=hi)
    assert torch.equal(o2[live], sep[live]) and torch.equal(o2[rows], hi[rows])


def test_gemm_bf16_split_stream_is_repeatable_and_accumulates(sdp):
    """Twenty in-place residual GEMMs in a row on the split stream (what a deep model does to it): the error against
    an fp32 running sum stays at the 1e-5 level instead of growing like a bf16 stream's, and the run is bit-repeatable."""
    M, C, K = 2610, 768, 768
    A = rnd(M, K, seed=31, dtype=torch.bfloat16)
    W = rnd(C, K, seed=32, scale=0.3 / math.sqrt(K), dtype=torch.bfloat16)
    x = rnd(M, C, seed=33)
    delta = A.float() @ W.float().t()

    def run(split):
        hi, lo = _split(x)
        if not split:
            lo = None
        for _ in range(20):
            sdp.ops.gemm(A, W, hi, residual=hi, residual_lo=lo, out_lo=lo)
        return hi.float() + (lo.float() if split else 0.0), hi, lo

    ref = x + 20 * delta
    got, hi, lo = run(True)
    got2, hi2, lo2 = run(True)
    plain, _, _ = run(False)
    assert torch.equal(hi, hi2) and torch.equal(lo, lo2)
    e_split = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    e_plain = float((plain - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert e_split < 5e-5 and e_plain > 20 * e_split, (e_split, e_plain)


def test_split_stream_row_kernels(sdp):
    """fill_registers / pool_ln / tokens_to_nchw on the (hi, lo) planes."""
    B, S, R, C = 6, 21, 5, 96
    table = rnd(R, C, seed=23)
    x = rnd(B, S, C, seed=24)
    hi, lo = _split(x)
    sdp.ops.fill_registers(hi, table, act_lo=lo)
    full = hi.float() + lo.float()
    assert float((full[:, :R] - table.expand(B, R, C)).abs().max()) < 2.0 ** -15 * float(table.abs().max())
    assert torch.equal(full[:, R:], (x.to(torch.bfloat16).float() + (x - x.to(torch.bfloat16).float()).to(torch.bfloat16).float())[:, R:])
    w, b = rnd(C, seed=25) + 1, rnd(C, seed=26)
    out = torch.empty(B, C, device="cuda", dtype=torch.float32)
    sdp.ops.pool_ln(hi, 0, R, w, b, out, 1e-5, act_lo=lo)
    ref = F.layer_norm(full[:, :R].mean(1), (C,), w, b, 1e-5)
    assert float((out - ref).abs().max()) < 2e-5
    T = S - R
    xo, ro = torch.empty(B, C, 4, 4, device="cuda"), torch.empty(B, R, C, device="cuda")
    sdp.ops.tokens_to_nchw(hi, xo, ro, T, R, act_lo=lo)
    assert torch.equal(xo.flatten(2).transpose(1, 2), full[:, R:]) and torch.equal(ro, full[:, :R])


def test_gemm_bf16_patch_embed_split_stream(sdp):
    """Patch-embedding epilogue (row remap + fp32 position table) writing both planes of the split stream."""
    B, T, R, C, K = 5, 36, 4, 128, 200
    S = T + R
    A = rnd(B * T, K, seed=41, dtype=torch.bfloat16)
    W = rnd(C, K, seed=42, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    pos = rnd(T, C, seed=43)
    hi = torch.zeros(B * S, C, device="cuda", dtype=torch.bfloat16)
    lo = torch.zeros_like(hi)
    sdp.ops.gemm(A, W, hi, residual=pos, res_first=True, res_mod=T, seq_remap=(T, S, R), out_lo=lo)
    ref = (A.float() @ W.float().t()).view(B, T, C) + pos
    got = (hi.float() + lo.float()).view(B, S, C)
    assert float((got[:, R:] - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert bool((got[:, :R] == 0).all())

@pytest.mark.parametrize("M,N,K", [(70, 50, 33), (129, 65, 100), (300, 128, 64)])
def test_gemm_fp32(sdp, M, N, K):
    A, W = rnd(M, K, seed=13), rnd(N, K, seed=14, scale=1 / math.sqrt(K))
    bias, res = rnd(N, seed=15), rnd(M, N, seed=16)
    out = torch.empty(M, N, device="cuda")
    sdp.ops.gemm(A, W, out, bias=bias, act="gelu", residual=res)
    ref = F.gelu(A.double() @ W.double().t() + bias.double()) + res.double()
    assert (out.double() - ref).abs().max() < 2e-5


def test_gemm_rejects_misaligned_pitch(sdp):
    A = rnd(16, 36, dtype=torch.bfloat16)     # 72-byte rows: not a legal TMA pitch
    W = rnd(16, 36, dtype=torch.bfloat16)
    with pytest.raises(sdp._lib.SdpNetLibraryError):
        sdp.ops.gemm(A, W, torch.empty(16, 16, device="cuda", dtype=torch.bfloat16))


# --------------------------------------------------------------------------------------------
# activations (epilogue functions), incl. the fast erf-GELU used by the bf16 epilogue
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["relu", "gelu", "gelu_tanh", "tanh", "sigmoid", "leaky_relu", "selu", "kelu"])
def test_activation_fp32_exact(sdp, act):
    x = torch.linspace(-8, 8, 20001, device="cuda")
    y = sdp.ops.activation(x, act)
    assert (y - O.ACTIVATIONS[act](x)).abs().max() < 2e-6


def test_fast_gelu_close_to_erf_gelu(sdp):
    x = torch.linspace(-10, 10, 200001, device="cuda")
    y = sdp.ops.activation(x, "gelu", force_fast=True)
    ref = F.gelu(x.double()).float()
    assert (y - ref).abs().max() < 2e-6      # << bf16 resolution; documented in DESIGN.md


# --------------------------------------------------------------------------------------------
# LayerNorm rows / pooled head front / register fill / layout bridges / im2col
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C", [(1, 8), (37, 32), (1000, 128), (523, 512), (2610, 768), (33, 1024), (20, 36), (9, 2048),
                                 (5, 4096)])
def test_layernorm_rows(sdp, dtype, M, C):
    x = (rnd(M, C, seed=20) * 3 + 1.5).to(dtype)
    w, b = rnd(C, seed=21) + 1, rnd(C, seed=22)
    out = torch.empty_like(x)
    sdp.ops.layernorm_rows(x, w, b, out, 1e-5)
    ref = F.layer_norm(x.float(), (C,), w, b, 1e-5)
    # bf16: half an output ulp (2^-9 relative) plus the input's own rounding through the affine
    tol = 2 ** -7 * float(ref.abs().max()) if dtype == torch.bfloat16 else 2e-5
    assert (out.float() - ref).abs().max() < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_ln_and_fill_registers(sdp, dtype):
    B, S, R, C = 6, 21, 5, 96
    table = rnd(R, C, seed=23)
    act = rnd(B, S, C, seed=24).to(dtype)
    sdp.ops.fill_registers(act, table)
    assert torch.allclose(act[:, :R].float(), table.to(dtype).float().expand(B, R, C))
    w, b = rnd(C, seed=25) + 1, rnd(C, seed=26)
    out = torch.empty(B, C, device="cuda", dtype=dtype)
    sdp.ops.pool_ln(act, 0, R, w, b, out, 1e-5)
    ref = F.layer_norm(act[:, :R].float().mean(1), (C,), w, b, 1e-5)
    assert (out.float() - ref).abs().max() < (3e-2 if dtype == torch.bfloat16 else 2e-5)
    out2 = torch.empty(B, C, device="cuda", dtype=torch.float32)
    sdp.ops.pool_ln(act, R, S - R, None, None, out2, 0.0)
    assert (out2 - act[:, R:].float().mean(1)).abs().max() < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_token_bridges_roundtrip(sdp, dtype):
    B, C, Gh, Gw, R = 3, 40, 5, 7, 4
    x, reg = rnd(B, C, Gh, Gw, seed=27), rnd(B, R, C, seed=28)
    act = torch.empty(B, R + Gh * Gw, C, device="cuda", dtype=dtype)
    sdp.ops.tokens_from_nchw(x, reg, act)
    ref = torch.cat([reg, x.flatten(2).transpose(1, 2)], 1)
    assert torch.equal(act.float(), ref.to(dtype).float())
    x2, r2 = torch.empty_like(x), torch.empty_like(reg)
    sdp.ops.tokens_to_nchw(act, x2, r2, Gh * Gw, R)
    assert torch.equal(x2, x.to(dtype).float()) and torch.equal(r2, reg.to(dtype).float())


@pytest.mark.parametrize("p,H,W", [(14, 28, 42), (16, 32, 32), (2, 8, 6), (4, 16, 16)])
def test_im2col_matches_conv(sdp, p, H, W):
    B, C = 3, 24
    x = rnd(B, 3, H, W, seed=29)
    w = rnd(C, 3, p, p, seed=30, scale=0.1)
    Kc = 3 * p * p
    Kp = (Kc + 7) // 8 * 8
    A = torch.full((B * (H // p) * (W // p), Kp), float("nan"), device="cuda")
    sdp.ops.im2col_patches(x, A, p)
    assert (A[:, Kc:] == 0).all()
    y = (A[:, :Kc] @ w.reshape(C, Kc).t()).view(B, H // p, W // p, C).permute(0, 3, 1, 2)
    assert (y - F.conv2d(x, w, stride=p)).abs().max() < 1e-4


@pytest.mark.parametrize("xd,ad", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("p,H,W,B", [(14, 224, 224, 5), (16, 224, 224, 3), (14, 28, 42, 2), (7, 21, 35, 2), (3, 9, 6, 4), (2, 8, 6, 3)])
def test_im2col_exact_in_every_dtype(sdp, xd, ad, p, H, W, B):
    """Pure data movement: every element equals the unfolded image converted once to the output dtype (even patch
    sizes take the pair kernel, odd ones the scalar kernel), padding columns are zero."""
    x = rnd(B, 3, H, W, seed=33).to(xd)
    Kc = 3 * p * p
    Kp = (Kc + 7) // 8 * 8
    A = torch.full((B * (H // p) * (W // p), Kp), float("nan"), device="cuda", dtype=ad)
    sdp.ops.im2col_patches(x, A, p)
    ref = F.unfold(x.float(), p, stride=p).transpose(1, 2).reshape(-1, Kc).to(ad)
    assert torch.equal(A[:, :Kc], ref) and bool((A[:, Kc:] == 0).all())


def test_embed_tokens(sdp):
    B, T, R, C = 3, 12, 2, 16
    act = rnd(B, R + T, C, seed=31)
    pos = rnd(T, C, seed=32)
    ref = act.clone()
    ref[:, R:] = F.gelu(ref[:, R:] + pos)
    sdp.ops.embed_tokens(act, pos, R, "gelu")
    assert (act - ref).abs().max() < 2e-6


# --------------------------------------------------------------------------------------------
# channel LayerNorm + depthwise conv
# --------------------------------------------------------------------------------------------
def dw_ref(act, R, Gh, Gw, gamma, beta, wdw, bdw, eps=1e-6):
    B, S, C = act.shape
    k = wdw.shape[-1]
    x = act[:, R:].float()
    xn = F.layer_norm(x, (C,), gamma, beta, eps)            # per-token LN over channels == layers.py:12-24
    img = xn.transpose(1, 2).reshape(B, C, Gh, Gw)
    lo = (k - 1) // 2
    img = F.pad(img, (lo, k - 1 - lo, lo, k - 1 - lo))      # zeros AFTER the norm
    y = F.conv2d(img, wdw.view(C, 1, k, k), bdw, groups=C)
    out = torch.zeros(B, S, C, device=act.device)
    out[:, R:] = y.flatten(2).transpose(1, 2)
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Gh,Gw,C,k,R,bias", [(16, 16, 768, 7, 5, False), (14, 14, 96, 7, 4, True), (8, 8, 32, 5, 1, True),
                                              (4, 6, 40, 3, 2, False), (5, 3, 16, 9, 3, True), (2, 2, 128, 7, 5, False)])
def test_ln_dwconv(sdp, dtype, Gh, Gw, C, k, R, bias):
    B = 3
    act = (rnd(B, R + Gh * Gw, C, seed=40) * 2 + 0.3).to(dtype)
    gamma, beta = rnd(C, seed=41) * 0.3 + 1, rnd(C, seed=42) * 0.3
    wdw = rnd(C, k, k, seed=43, scale=1 / k)
    bdw = rnd(C, seed=44) if bias else None
    out = torch.full_like(act, float("nan"))
    sdp.ops.ln_dwconv(act, gamma, beta, wdw.reshape(C, k * k).t().contiguous(), bdw, out, Gh, Gw, R)
    ref = dw_ref(act, R, Gh, Gw, gamma, beta, wdw, bdw)
    assert (out[:, :R] == 0).all()
    tol = 4e-2 if dtype == torch.bfloat16 else 5e-5
    assert (out.float() - ref).abs().max() < tol
    if dtype == torch.bfloat16 and k in (3, 5, 7) and C % 8 == 0:
        # token statistics supplied by a producer instead of recomputed in the kernel
        stats = torch.empty(B * (R + Gh * Gw), 4, 2, device="cuda")
        sdp.ops.row_stats(act, stats)
        out2 = torch.full_like(act, float("nan"))
        sdp.ops.ln_dwconv(act, gamma, beta, wdw.reshape(C, k * k).t().contiguous(), bdw, out2, Gh, Gw, R, stats=stats)
        assert (out2.float() - ref).abs().max() < tol


@pytest.mark.parametrize("B,Gh,Gw,C,k,R,bias", [(3, 16, 16, 768, 7, 5, False), (5, 16, 16, 64, 7, 4, True), (2, 8, 8, 32, 5, 1, True),
                                                (7, 12, 16, 96, 3, 0, False), (40, 5, 8, 32, 7, 2, True), (1, 16, 8, 128, 5, 5, False),
                                                # grids narrower than the 16-slot rows of the shared-memory tiles
                                                (3, 14, 14, 768, 7, 5, False), (4, 14, 14, 512, 7, 5, True), (3, 6, 3, 32, 3, 1, True),
                                                (2, 4, 15, 64, 5, 2, False), (5, 2, 1, 32, 7, 3, True), (2, 10, 9, 96, 7, 0, False)])
def test_ln_dwconv_slab(sdp, B, Gh, Gw, C, k, R, bias):
    """Channel-stationary tensor-core kernel (dwconv_slab.cu) against the same torch reference (layers.py:102)."""
    assert sdp.ops.ln_dwconv_slab_ok(Gh, Gw, C, k, torch.bfloat16)
    assert not sdp.ops.ln_dwconv_slab_ok(15, 15, C, k, torch.bfloat16)      # odd token count
    assert not sdp.ops.ln_dwconv_slab_ok(17, 16, C, k, torch.bfloat16)
    act = (rnd(B, R + Gh * Gw, C, seed=45) * 2 + 0.3).bfloat16()
    gamma, beta = rnd(C, seed=41) * 0.3 + 1, rnd(C, seed=42) * 0.3
    wdw = rnd(C, k, k, seed=43, scale=1 / k)
    bdw = rnd(C, seed=44) if bias else None
    out = torch.full_like(act, float("nan"))
    scratch = torch.full((2 * B * Gh * Gw,), float("nan"), device="cuda")
    sdp.ops.ln_dwconv_slab(act, scratch, gamma, beta, wdw.reshape(C, k * k).t().contiguous(), bdw, out, Gh, Gw, R)
    ref = dw_ref(act, R, Gh, Gw, gamma, beta, wdw, bdw)
    assert (out[:, :R] == 0).all()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() < 4e-2
    # statistics in the producer-GEMM format: (sum, sumsq) column parts per row of the [B*S, C] activation
    for parts in (2, 6):
        stats = torch.zeros(B * (R + Gh * Gw), parts, 2, device="cuda")
        xf = act.float().view(-1, C)
        for p_, chunk in enumerate(xf.chunk(parts, dim=1)):
            stats[:, p_, 0] = chunk.sum(1)
            stats[:, p_, 1] = (chunk * chunk).sum(1)
        out2 = torch.full_like(act, float("nan"))
        scratch.fill_(float("nan"))
        sdp.ops.ln_dwconv_slab(act, scratch, gamma, beta, wdw.reshape(C, k * k).t().contiguous(), bdw, out2, Gh, Gw, R,
                               producer_stats=stats)
        assert (out2[:, :R] == 0).all()
        assert (out2.float() - ref).abs().max() < 4e-2


# --------------------------------------------------------------------------------------------
# attention with fused QK LayerNorm
# --------------------------------------------------------------------------------------------
def attn_ref(qkv, h, qn_w, qn_b, kn_w, kn_b, eps=1e-5):
    B, S, C3 = qkv.shape
    C = C3 // 3
    d = C // h
    q, k, v = [t.float().view(B, S, h, d).transpose(1, 2) for t in qkv.split(C, dim=-1)]
    if qn_w is not None:
        q = F.layer_norm(q, (d,), qn_w, qn_b, eps)
        k = F.layer_norm(k, (d,), kn_w, kn_b, eps)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1)
    return (p @ v).transpose(1, 2).reshape(B, S, C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("S,h,d,norm", [(261, 8, 96, True), (201, 8, 64, True), (201, 8, 96, False), (9, 4, 32, True),
                                        (68, 2, 16, True), (20, 4, 8, True), (33, 3, 24, False), (300, 2, 128, True),
                                        (17, 4, 32, True), (32, 2, 64, True), (261, 8, 96, False), (201, 8, 64, False),
                                        (513, 2, 96, False), (400, 2, 64, False), (256, 4, 32, False),
                                        (16, 2, 16, False), (300, 1, 128, False), (1, 2, 32, False),
                                        # shapes of the tcgen05 kernel: one / two / three query tiles, 16-key tails,
                                        # one and two TMA boxes of keys, all three head sizes
                                        (128, 2, 64, False), (5, 2, 64, False), (16, 1, 64, False), (257, 2, 96, False),
                                        (288, 2, 64, False), (272, 1, 128, False), (256, 2, 128, False),
                                        (144, 3, 96, False), (208, 2, 128, False)])
def test_attention(sdp, dtype, S, h, d, norm):
    B, C = 2, h * d
    qkv = rnd(B, S, 3 * C, seed=50, scale=1.5).to(dtype)
    if norm:
        qn_w, qn_b = rnd(d, seed=51) * 0.3 + 1, rnd(d, seed=52) * 0.3
        kn_w, kn_b = rnd(d, seed=53) * 0.3 + 1, rnd(d, seed=54) * 0.3
    else:
        qn_w = qn_b = kn_w = kn_b = None
    out = torch.full((B, S, C), float("nan"), device="cuda", dtype=dtype)
    sdp.ops.attention(qkv, out, h, qn_w, qn_b, kn_w, kn_b)
    ref = attn_ref(qkv, h, qn_w, qn_b, kn_w, kn_b)
    assert torch.isfinite(out.float()).all()
    tol = 3e-2 if dtype == torch.bfloat16 else 5e-5
    assert (out.float() - ref).abs().max() < tol


@pytest.mark.parametrize("S,h,d", [(261, 8, 96), (257, 2, 96), (201, 8, 64), (133, 2, 128)])
def test_attention_many_ctas_repeatable(sdp, S, h, d):
    """Several waves of CTAs on warm caches, four runs on identical inputs: bit-identical outputs and correct against
    the fp32 reference.  (A missing barrier behind the transposed tail tile once passed every two-image test and
    corrupted the last rows of most images at batch 1024.)"""
    B, C = 1200 // h * 2, h * d
    qkv = rnd(B, S, 3 * C, seed=60).to(torch.bfloat16)
    outs = []
    for _ in range(4):
        o = torch.full((B, S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        sdp.ops.attention(qkv, o, h)
        outs.append(o)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = attn_ref(qkv[:24], h, None, None, None, None)
    for o in (outs[0], outs[3]):
        assert (o[:24].float() - ref).abs().max() < 2e-2
        assert (o[-24:].float() - attn_ref(qkv[-24:], h, None, None, None, None)).abs().max() < 2e-2


def _score_bound(qkv, h):
    """max |q| max |k| / sqrt(d): a valid bound of every score of the batch (nats)."""
    B, S, C3 = qkv.shape
    C = C3 // 3
    d = C // h
    q, k = qkv[..., :C].float().view(B, S, h, d), qkv[..., C:2 * C].float().view(B, S, h, d)
    return float(q.norm(dim=-1).max() * k.norm(dim=-1).max() / math.sqrt(d)) * 1.001


@pytest.mark.parametrize("S,h,d", [(261, 8, 96), (201, 8, 96), (201, 8, 64), (257, 2, 96), (133, 2, 128), (128, 2, 64), (5, 2, 64),
                                   (16, 1, 64), (288, 2, 64), (272, 1, 128), (144, 3, 96), (208, 2, 128)])
def test_attention_bounded_one_pass(sdp, S, h, d):
    """sdp_attention_bounded: with a bound on |q k^T| / sqrt(d) the tcgen05 kernel takes exp(s) without a row maximum
    (softmax is invariant to the exponent reference) -- same result as the two-pass kernel and the fp32 reference."""
    B, C = 3, h * d
    qkv = (rnd(B, S, 3 * C, seed=61) * 0.6).to(torch.bfloat16)
    sb = _score_bound(qkv, h)
    assert 0 < sb < 60
    two = torch.full((B, S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    one = torch.full_like(two, float("nan"))
    sdp.ops.attention(qkv, two, h)
    sdp.ops.attention(qkv, one, h, score_bound=sb)
    ref = attn_ref(qkv, h, None, None, None, None)
    assert torch.isfinite(one.float()).all()
    assert (one.float() - ref).abs().max() < 2e-2
    assert (one.float() - two.float()).abs().max() < 2e-2
    # a bound past the one-pass limit (or none) is the two-pass kernel, bit for bit
    far = torch.full_like(two, float("nan"))
    sdp.ops.attention(qkv, far, h, score_bound=75.0)
    assert torch.equal(far, two)


def test_attention_bounded_extreme_rows(sdp):
    """Rows whose scores sit at the ends of the bound: every score of a row at +bound or at -bound (exp(s) at e^+-40),
    and a row with one dominant key.  Finite, and equal to the reference."""
    S, h, d = 261, 8, 96
    B, C = 2, h * d
    g = _g(62)
    k = torch.randn(B, S, h, d, generator=g, device="cuda")
    k = k / k.norm(dim=-1, keepdim=True)
    q = torch.randn(B, S, h, d, generator=g, device="cuda")
    q = q / q.norm(dim=-1, keepdim=True)
    amp = 40.0 * math.sqrt(d)                               # |q| |k| / sqrt(d) = 40 nats
    q[:, 0] = k[:, 7]                                       # one dominant key: score +40 against ~0
    q[:, 1] = -k[:, 9]                                      # one key at -40
    k[:, 100:] = k[:, 100:101]                              # many identical keys ...
    q[:, 2] = k[:, 100]                                     # ... all at +40 for this row
    q[:, 3] = -k[:, 100]                                    # ... and at -40 for this one
    v = torch.randn(B, S, h, d, generator=g, device="cuda")
    qkv = torch.cat([(q * amp).reshape(B, S, C), k.reshape(B, S, C), v.reshape(B, S, C)], -1).to(torch.bfloat16).contiguous()
    sb = _score_bound(qkv, h)
    assert 39 < sb < 42
    out = torch.full((B, S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    sdp.ops.attention(qkv, out, h, score_bound=sb)
    ref = attn_ref(qkv, h, None, None, None, None)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() < 3e-2


@pytest.mark.parametrize("S,h,d", [(261, 8, 96), (201, 8, 64), (133, 2, 128)])
def test_attention_bounded_many_ctas_repeatable(sdp, S, h, d):
    B, C = 1200 // h * 2, h * d
    qkv = (rnd(B, S, 3 * C, seed=63) * 0.6).to(torch.bfloat16)
    sb = _score_bound(qkv, h)
    outs = []
    for _ in range(3):
        o = torch.full((B, S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        sdp.ops.attention(qkv, o, h, score_bound=sb)
        outs.append(o)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    for sl in (slice(0, 24), slice(B - 24, B)):
        assert (outs[0][sl].float() - attn_ref(qkv[sl], h, None, None, None, None)).abs().max() < 2e-2


def test_qk_score_bound_holds_for_layernormed_heads(sdp):
    """ops.qk_score_bound (what the engine hands to sdp_attention_bounded): no score of LayerNorm-ed, bf16-rounded
    q and k exceeds it (layers.py:286,289-291)."""
    d, n = 96, 4096
    g = _g(64)
    for scale, shift in ((1.0, 0.0), (1.7, 0.4), (0.3, 2.0)):
        qw = 1 + 0.3 * torch.randn(d, generator=g, device="cuda") * scale
        kw = 1 + 0.3 * torch.randn(d, generator=g, device="cuda") * scale
        qb, kb = shift * torch.randn(d, generator=g, device="cuda"), shift * torch.randn(d, generator=g, device="cuda")
        x, y = torch.randn(n, d, generator=g, device="cuda") * 5 + 3, torch.randn(n, d, generator=g, device="cuda")
        y[: n // 2] = x[: n // 2]                           # aligned pairs: where the bound is approached
        q = F.layer_norm(x, (d,), qw, qb, 1e-5).bfloat16().float()
        k = F.layer_norm(y, (d,), kw, kb, 1e-5).bfloat16().float()
        smax = float((q @ k.t()).abs().max()) / math.sqrt(d)
        assert smax <= sdp.ops.qk_score_bound(qw, qb, kw, kb)


@pytest.mark.parametrize("B,K,ls", [(37, 100, 0.0), (512, 1000, 0.1), (3, 7, 0.0)])
def test_eval_metrics_kernel(sdp, B, K, ls):
    """On-device CE / BCE / top-1 accumulation (model_test.py:76-85) against the oracle's definition."""
    logits = rnd(B, K, seed=90, scale=3.0)
    labels = torch.randint(0, K, (B,), generator=_g(91), device="cuda")
    logits[torch.arange(0, B, 3, device="cuda"), labels[::3]] += 8.0
    logits[1, :] = 0.25                                      # all-equal row: argmax must be index 0 like torch
    meter = sdp.evaluate.EvalMeter("cuda", ls)
    meter.update(logits[: B // 2], labels[: B // 2])         # two batches accumulate
    meter.update(logits[B // 2:], labels[B // 2:])
    got = meter.result()
    ref = O.eval_metrics(logits.cpu(), labels.cpu(), ls)
    assert got["samples"] == B
    assert abs(got["cross_entropy"] - ref["cross_entropy"]) < 1e-4
    assert abs(got["bce_with_logits"] - ref["bce_with_logits"]) < 1e-5
    assert abs(got["accuracy"] - ref["accuracy"]) < 1e-9
    # a running readout does not disturb the accumulator (result() reduces a clone): same numbers twice
    assert meter.result() == got


def test_eval_metrics_rejects_labels_out_of_range(sdp):
    """nn.CrossEntropyLoss semantics (model_test.py:66,80): ignore_index -100 rows are skipped, any other label
    outside [0, K) is an error -- never an out-of-bounds read folded into the sums."""
    B, K = 64, 10
    logits = rnd(B, K, seed=92)
    labels = torch.randint(0, K, (B,), generator=_g(93), device="cuda")
    ign = labels.clone()
    ign[::4] = -100
    meter = sdp.evaluate.EvalMeter("cuda")
    meter.update(logits, ign)
    got = meter.result()
    keep = ign >= 0
    ref = O.eval_metrics(logits[keep].cpu(), labels[keep].cpu(), 0.0)
    assert got["samples"] == int(keep.sum()) and abs(got["cross_entropy"] - ref["cross_entropy"]) < 1e-4
    for bad in (-1, K, 10 ** 6):
        lb = labels.clone()
        lb[5] = bad
        meter = sdp.evaluate.EvalMeter("cuda")
        meter.update(logits, lb)
        with pytest.raises(ValueError):
            meter.result()


def test_ops_run_on_the_tensors_device_not_the_current_one(sdp):
    """Every op launches on its tensors' device and that device's current stream (a model on cuda:1 while cuda:0 is
    current); tensors on different devices are refused.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with torch.cuda.device(0):
        A = torch.randn(300, 256, device="cuda:1").bfloat16()
        W = torch.randn(512, 256, device="cuda:1").bfloat16()
        out = torch.empty(300, 512, device="cuda:1", dtype=torch.bfloat16)
        sdp.ops.gemm(A, W, out, act="gelu")
        torch.cuda.synchronize(1)
        ref = torch.nn.functional.gelu(A.float() @ W.float().t())
        assert relerr(out, ref) < 1.5e-2
        with pytest.raises(RuntimeError):
            sdp.ops.gemm(A, W.to("cuda:0"), out)


def test_torch_ops_cpp_wrappers_match_the_ctypes_path(sdp):
    """torch.ops.sdpnet_b200.* (TORCH_LIBRARY wrappers, csrc/torch_ops.cpp) give the same bits as `sdp.ops.*`."""
    t = torch.ops.sdpnet_b200
    A = rnd(300, 256, seed=1, dtype=torch.bfloat16)
    W = rnd(512, 256, seed=2, scale=1 / 16, dtype=torch.bfloat16)
    bias, res = rnd(512, seed=3), rnd(300, 512, seed=4, dtype=torch.bfloat16)
    o1, o2 = torch.empty(300, 512, device="cuda", dtype=torch.bfloat16), torch.empty(300, 512, device="cuda", dtype=torch.bfloat16)
    t.gemm(A, W, o1, bias, res, sdp.ops.act_id("gelu"))
    sdp.ops.gemm(A, W, o2, bias=bias, residual=res, act="gelu")
    assert torch.equal(o1, o2)
    x, w, b = rnd(77, 96, seed=5, dtype=torch.bfloat16), rnd(96, seed=6) + 1, rnd(96, seed=7)
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    t.layernorm_rows(x, w, b, y1, 1e-5)
    sdp.ops.layernorm_rows(x, w, b, y2, 1e-5)
    assert torch.equal(y1, y2)
    B, Gh, Gw, C, k, R = 3, 8, 8, 64, 5, 2
    act = rnd(B, R + Gh * Gw, C, seed=8, dtype=torch.bfloat16)
    g_, be, wd = rnd(C, seed=9) + 1, rnd(C, seed=10), rnd(k * k, C, seed=11, scale=0.2)
    d1, d2 = torch.empty_like(act), torch.empty_like(act)
    t.ln_dwconv(act, g_, be, wd, None, d1, Gh, Gw, R, 1e-6)
    sdp.ops.ln_dwconv(act, g_, be, wd, None, d2, Gh, Gw, R, 1e-6)
    assert torch.equal(d1, d2)
    qkv = rnd(2, 133, 3 * 128, seed=12, dtype=torch.bfloat16)
    a1, a2 = torch.empty(2, 133, 128, device="cuda", dtype=torch.bfloat16), torch.empty(2, 133, 128, device="cuda", dtype=torch.bfloat16)
    t.attention(qkv, a1, 2, None, None, None, None, 1e-5)
    sdp.ops.attention(qkv, a2, 2)
    assert torch.equal(a1, a2)
    with pytest.raises(RuntimeError):
        t.gemm(A, W[:, :100].contiguous(), o1, None, None, 0)          # K mismatch: TORCH_CHECK


def test_launch_counter(sdp):
    sdp.ops.launch_count(reset=True)
    x = rnd(4, 64)
    sdp.ops.layernorm_rows(x, None, None, torch.empty_like(x), 1e-5)
    assert sdp.ops.launch_count() == 1
