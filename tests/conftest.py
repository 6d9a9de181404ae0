import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# PyTorch fp32 references must really be fp32 on the GPU (no TF32 in conv / matmul)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_fixture(name):
    """-> (meta dict, arrays dict of torch tensors, state_dict or None)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    arrs = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta" and not k.startswith("sd/")}
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return meta, arrs, (sd or None)


def model_fixture_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN)
                  if f.endswith(".npz") and not f.startswith(("layer_", "act_", "preprocess_", "ckpt_")))


def fixture_inputs(meta, embedded_sd=None):
    """Rebuild the weights and input a fixture was generated from (oracle/make_golden.py)."""
    import sdpnet_oracle as O
    sd = embedded_sd if embedded_sd is not None else \
        O.synth_state_dict(meta["cfg"], seed=meta["seed"], stress=meta["stress"])
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    if abs(chk - meta["checksum"]) > 1e-6 * max(1.0, abs(meta["checksum"])):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    x = torch.randn(meta["B"], 3, meta["H"], meta["W"],
                    generator=torch.Generator().manual_seed(meta["input_seed"]))
    return sd, x
