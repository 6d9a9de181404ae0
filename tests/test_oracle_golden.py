"""Pins oracle/sdpnet_oracle.py against the reference-generated golden vectors (CPU)."""
import numpy as np
import pytest
import torch

import sdpnet_oracle as O
from conftest import GOLDEN, fixture_inputs, load_fixture, model_fixture_names

TOL = 2e-5  # fp32 restatement vs fp32 reference on O(1..10) activations (summation order)


@pytest.mark.parametrize("name", model_fixture_names())
def test_oracle_matches_reference_fixture(name):
    meta, arrs, sd_emb = load_fixture(name)
    sd, x = fixture_inputs(meta, sd_emb)
    stages = {}
    logits, x_raw, reg = O.forward(sd, meta["cfg"], x, meta["num_registers"], True, stages=stages)
    assert logits.shape == arrs["logits"].shape
    assert (logits - arrs["logits"]).abs().max() < TOL
    assert (x_raw - arrs["x_raw"]).abs().max() < TOL
    assert (reg - arrs["registers"]).abs().max() < TOL
    n = 0
    for k, v in arrs.items():
        if k.startswith("stage/"):
            assert (stages[k[6:]] - v).abs().max() < TOL, k
            n += 1
    assert n >= 4


def test_oracle_fp64_close_to_fp32():
    meta, arrs, sd_emb = load_fixture("yaml_r5_stress")
    sd, x = fixture_inputs(meta, sd_emb)
    l64 = O.forward(sd, meta["cfg"], x, meta["num_registers"], dtype=torch.float64)
    assert (l64.float() - arrs["logits"]).abs().max() < TOL


def test_kelu_matches_reference_grid():
    z = np.load(f"{GOLDEN}/act_kelu_grid.npz")
    y = O.kelu(torch.from_numpy(z["x"]))
    assert (y - torch.from_numpy(z["y"])).abs().max() < 1e-6


def test_encoder_layer_kelu_fixture():
    z = np.load(f"{GOLDEN}/layer_encoder_kelu.npz")
    sd = {"t." + k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    xo, ro = O.encoder_layer(torch.from_numpy(z["x"]), torch.from_numpy(z["reg"]), sd, "t.", 4, "kelu")
    assert (xo - torch.from_numpy(z["x_out"])).abs().max() < TOL
    assert (ro - torch.from_numpy(z["reg_out"])).abs().max() < TOL


def test_conv_embedding_register_limit():
    # conv_embedding cannot serve R = max_num_registers (IndexError upstream)
    cfg = dict(embedding_dim=16, n_head=2, num_blocks=1, patch_size=4, output_classes=4,
               max_image_size=[4, 4], conv_embedding=True, max_num_registers=5)
    sd = O.synth_state_dict(cfg, 0)
    with pytest.raises(IndexError):
        O.forward(sd, cfg, torch.randn(1, 3, 16, 16), num_registers=4)


def test_flops_formula_matches_baseline_table():
    yaml_cfg = dict(n_head=8, conv_kernel_size=7, conv_block_num=2, ff_multiplication_factor=4,
                    head_output_from_register=True, simple_mlp_output=False, output_classes=1000)
    xl = dict(yaml_cfg, embedding_dim=768, num_blocks=17, patch_size=14)
    m = dict(yaml_cfg, embedding_dim=768, num_blocks=12, patch_size=16)
    s = dict(yaml_cfg, embedding_dim=512, num_blocks=12, patch_size=16)
    assert abs(O.flops_per_image(xl, 224, 224, 5) / 1e9 - 163.569) < 0.01
    assert abs(O.flops_per_image(m, 224, 224, 5) / 1e9 - 89.133) < 0.01
    assert abs(O.flops_per_image(s, 224, 224, 5) / 1e9 - 40.105) < 0.01
