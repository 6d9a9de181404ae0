"""CPU tests of the 'next' rows: checkpoint loaders (host logic) and the oracle of the
evaluation metrics against reference-generated golden values."""
import io
import os

import numpy as np
import pytest
import torch

import sdpnet_oracle as O
from conftest import GOLDEN


@pytest.fixture(scope="module")
def sdp():
    import sdpnet_b200 as m
    return m


CFG = dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4, output_classes=10, max_image_size=[4, 4],
           head_output_from_register=True, conv_first=False)


def _save(obj):
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return buf


def test_trainer_snapshot_with_ddp_and_compile_prefixes(sdp):
    sd = O.synth_state_dict(CFG, seed=3)
    ddp = {"module._orig_mod." + k: v for k, v in sd.items()}          # DDP(compile(model)) key shape
    blob = {"model_state_dict": ddp, "model_config": CFG, "optimizer_state": {}, "scheduler_state": {}, "epoch": 17}
    cfg, got, extras = sdp.checkpoint.read_checkpoint(_save(blob))      # training_tools.py:214-219
    assert cfg == CFG and extras == {"epoch": 17}
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    model = sdp.checkpoint.load_model(_save(blob), device=None)         # CPU: construction + strict load only
    assert model.config == CFG and not model.training
    assert all(torch.equal(model.state_dict()[k], sd[k]) for k in sd)


def test_checkpoints_written_by_the_reference_itself(sdp):
    """Files produced by the reference's own writers (oracle/make_golden_checkpoints.py: SdPModel.save_model,
    Trainer._save_checkpoint behind a DDP-style `module.` wrapper, EMA_model.save_ema_model): layouts recognised,
    prefixes stripped, strict load into the mirrored modules, from_pretrained reads the save_model file."""
    paths = {k: os.path.join(GOLDEN, f"ckpt_{k}.pt") for k in ("save_model", "trainer", "ema")}
    cfg_a, sd_a, ex_a = sdp.checkpoint.read_checkpoint(paths["save_model"])
    cfg_b, sd_b, ex_b = sdp.checkpoint.read_checkpoint(paths["trainer"])
    cfg_c, sd_c, _ = sdp.checkpoint.read_checkpoint(paths["ema"])
    assert cfg_a == cfg_b and cfg_c is None and ex_a == {} and ex_b == {"epoch": 7}
    raw = torch.load(paths["trainer"], map_location="cpu")
    assert all(k.startswith("module.") for k in raw["model_state_dict"])          # what the reference really wrote
    assert set(sd_a) == set(sd_b) == set(sd_c) and not any(k.startswith("module.") for k in sd_b)
    assert all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a)
    model, ema = sdp.checkpoint.load_model(paths["trainer"], ema_path=paths["ema"], device=None)
    assert model.config == cfg_a and not model.training
    assert all(torch.equal(model.state_dict()[k], sd_a[k]) for k in sd_a)
    k = "blocks.1.conv_blocks.0.conv1d.0.weight"
    assert not torch.equal(ema.state_dict()[k], model.state_dict()[k])
    again = sdp.MainModel.from_pretrained(paths["save_model"])
    assert again.config == cfg_a and all(torch.equal(again.state_dict()[k], sd_a[k]) for k in sd_a)
    # and the oracle reproduces the reference's forward of those weights (the fixture the GPU test checks against)
    z = np.load(os.path.join(GOLDEN, "ckpt_expected.npz"))
    with torch.no_grad():
        lo, xr, rg = O.forward(sd_a, cfg_a, torch.from_numpy(z["x"]), 3, True)
    assert float((lo - torch.from_numpy(z["logits"])).abs().max()) < 2e-5
    assert float((xr - torch.from_numpy(z["x_raw"])).abs().max()) < 2e-4


def test_ema_file_and_save_model_roundtrip(sdp, tmp_path):
    sd = O.synth_state_dict(CFG, seed=4)
    ema = {"module." + k: v * 0.5 if v.is_floating_point() else v for k, v in sd.items()}   # training_tools.py:300-302
    cfg, got, _ = sdp.checkpoint.read_checkpoint(_save(ema))
    assert cfg is None and set(got) == set(sd)
    snap = {"model_state_dict": sd, "model_config": CFG, "optimizer_state": {}, "scheduler_state": {}, "epoch": 1}
    model, ema_model = sdp.checkpoint.load_model(_save(snap), ema_path=_save(ema), device=None)   # model_test.py:28-42
    k = "blocks.0.t_block.q_proj.weight"
    assert torch.allclose(ema_model.state_dict()[k], 0.5 * model.state_dict()[k])
    # SdPModel.save_model / from_pretrained (utility_layers.py:169-198)
    os.chdir(tmp_path)
    model.save_model("m")
    again = sdp.MainModel.from_pretrained("m.pt")
    assert again.config == CFG and all(torch.equal(again.state_dict()[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        sdp.checkpoint.read_checkpoint({"weights": 1, "other": "x"})
    with pytest.raises(ValueError):
        sdp.checkpoint.load_model(_save(ema), device=None)               # no config anywhere


def test_eval_metrics_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "act_eval_metrics.npz"))
    logits, labels = torch.from_numpy(z["logits"]), torch.from_numpy(z["labels"])
    m0 = O.eval_metrics(logits, labels, 0.0)
    m1 = O.eval_metrics(logits, labels, 0.1)
    assert abs(m0["cross_entropy"] - float(z["ce"])) < 1e-5
    assert abs(m0["accuracy"] - float(z["acc"])) < 1e-7
    assert abs(m0["bce_with_logits"] - float(z["bce_ls0.0"])) < 1e-6
    assert abs(m1["bce_with_logits"] - float(z["bce_ls0.1"])) < 1e-6
