"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes structs match the header's field order, and the host-side module API
(constructor surface, state_dict layout, error behaviour) needs no GPU.  No compute call is made."""
import ctypes
import os
import re
import sys

import pytest
import torch

from conftest import ROOT

import sdpnet_oracle as O

HEADER = os.path.join(ROOT, "include", "sdpnet_b200.h")


@pytest.fixture(scope="module")
def sdp():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "sdp-net_b200", "lib", "libsdpnet_b200.so")):
        g.build()
    import sdpnet_b200 as m
    return m


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sdp):
    lib = ctypes.CDLL(sdp._lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in sdp._lib.SYMBOLS, f"{n} has no ctypes prototype"
    assert sdp._lib.lib().sdp_abi_version() == int(re.search(r"SDPNET_B200_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_ctypes_structs_follow_header_field_order(sdp):
    src = open(HEADER).read()

    def fields(struct_name):
        end = src.index("} " + struct_name + ";")
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?(struct\s+)?\w+\s*", "", decl)
            out += [n.strip().lstrip("*").strip() for n in names.split(",")]
        return out

    L = sdp._lib
    assert fields("sdp_gemm_args") == [n for n, _ in L.GemmArgs._fields_]
    assert fields("sdp_encoder_weights") == [n for n, _ in L.EncoderWeights._fields_]
    assert fields("sdp_mixer_weights") == [n for n, _ in L.MixerWeights._fields_]
    assert fields("sdp_model_desc") == [n for n, _ in L.ModelDesc._fields_]
    assert fields("sdp_workspace") == [n for n, _ in L.Workspace._fields_]
    assert fields("sdp_image_desc") == [n for n, _ in L.ImageDesc._fields_] and ctypes.sizeof(L.ImageDesc) == 16


def test_torch_library_ops_are_registered_from_the_cpp_wrappers(sdp):
    """`torch.ops.sdpnet_b200.*` come from TORCH_LIBRARY in csrc/torch_ops.cpp (libsdpnet_b200_torch.so), CUDA-only:
    a CPU tensor has no kernel to dispatch to."""
    import torch
    sdp.ops.load_torch_ops()
    for name in ("gemm", "layernorm_rows", "ln_dwconv", "attention"):
        op = getattr(torch.ops.sdpnet_b200, name)
        assert "Tensor(a!) out" in str(op.default._schema)
    x = torch.zeros(4, 8)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.sdpnet_b200.layernorm_rows(x, None, None, torch.empty_like(x), 1e-5)


def test_no_cpu_fallback_and_loud_errors(sdp):
    cfg = dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4, output_classes=10, max_image_size=[4, 4])
    model = sdp.MainModel.from_dict(**cfg).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        sdp.ops.layernorm_rows(torch.randn(4, 8), None, None, torch.empty(4, 8), 1e-5)
    with pytest.raises(RuntimeError):
        sdp.Engine(cfg, model.state_dict(), "cpu")
    with pytest.raises(ValueError):
        sdp.ops.act_id("swish")


@pytest.mark.parametrize("name", ["yaml_r4_refinit", "cifar_path", "biases_relu_nonsquare", "headbias_mlp"])
def test_module_state_dict_layout_is_the_references(sdp, name):
    """Strict load of a reference-layout state_dict (the fixtures' key sets were validated against
    the real reference with load_state_dict(strict=True) when they were generated)."""
    from conftest import load_fixture
    meta, _, _ = load_fixture(name)
    sd = O.synth_state_dict(meta["cfg"], seed=meta["seed"], stress=meta["stress"])
    model = sdp.MainModel.from_dict(**meta["cfg"])
    model.load_state_dict(sd, strict=True)
    got = model.state_dict()
    assert set(got) == set(sd)
    assert all(got[k].shape == sd[k].shape and got[k].dtype == sd[k].dtype for k in sd)
    assert model.config == meta["cfg"]


def test_constructor_defaults_and_init_match_reference_contract(sdp):
    m = sdp.MainModel()          # reference defaults: model.py:28-54
    assert len(m.blocks) == 10 and m.blocks[0].conv_first is True
    assert m.conv_init.conv.weight.shape == (128, 3, 16, 16)
    w = m.blocks[0].t_block.q_proj.weight
    assert 0.008 < float(w.std()) < 0.012 and float(w.abs().max()) < 0.1   # trunc_normal_(std=0.01), cut at +-2
    assert m.return_num_params()["Trainable_params"] == sum(p.numel() for p in m.parameters())
    with pytest.raises(TypeError):
        sdp.MainModel(activation="kelu")
    with pytest.raises(ValueError):
        sdp.MainModel(activation="fast_gelu")
    with pytest.raises(AssertionError):
        sdp.StochasticDepth(0.0)


def test_qk_score_bound_is_an_upper_bound(sdp):
    """`ops.qk_score_bound` (handed to sdp_attention_bounded as the exponent range of the one-pass softmax): pure host
    arithmetic on the q_norm / k_norm parameters (layers.py:236-237,286); no score of LayerNorm-ed, bf16-rounded q and k
    may exceed it, aligned pairs come close."""
    import math
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    for d in (32, 64, 96):
        for scale, shift in ((0.0, 0.0), (1.0, 0.2), (2.0, 1.5)):
            qw, kw = 1 + 0.3 * scale * torch.randn(d, generator=g), 1 + 0.3 * scale * torch.randn(d, generator=g)
            qb, kb = shift * torch.randn(d, generator=g), shift * torch.randn(d, generator=g)
            x = torch.randn(2048, d, generator=g) * 4 + 1
            y = torch.randn(2048, d, generator=g)
            y[:1024] = x[:1024]
            q = F.layer_norm(x, (d,), qw, qb, 1e-5).bfloat16().float()
            k = F.layer_norm(y, (d,), kw, kb, 1e-5).bfloat16().float()
            smax = float((q @ k.t()).abs().max()) / math.sqrt(d)
            bound = sdp.ops.qk_score_bound(qw, qb, kw, kb)
            assert smax <= bound
            if scale == 0.0:
                assert bound == pytest.approx(1.02 * math.sqrt(d)) and smax > 0.9 * math.sqrt(d)
