"""TEST / BENCH INFRASTRUCTURE -- never imported by the product path (sdp-net_b200/).

Imports the UNMODIFIED reference (`MainModel`, /root/reference/model.py:27) from the git-ignored
`baseline/_ref/` copy made by `tools/vendor_reference.sh` (or straight from /root/reference in the
build container), so that the real reference can run next to the engine on the GPU box:

* `bench.py --impl reference`      -- its CPU forward on the host cores (cpu_baseline.kind = "reference")
* `bench.py` `gpu_comparators`     -- eager under `torch.autocast('cuda', bfloat16)` (training_tools.py:85)
                                      and `torch.compile(model)` in eval (model_test.py:16,64)
* `tests/test_gpu_reference.py`    -- engine vs reference, both on the GPU

`model.py:11` imports `training_utilities`, which imports `wandb` at its top (training_utilities.py:7);
only the import has to succeed, so a missing wandb is stubbed in `sys.modules`.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
_mod = None


def reference_dir():
    for d in _CANDIDATES:
        if os.path.exists(os.path.join(d, "model.py")) and os.path.exists(os.path.join(d, "layers.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


def load():
    """The reference's `model` module (cached).  Raises RuntimeError when no copy is present."""
    global _mod
    if _mod is not None:
        return _mod
    d = reference_dir()
    if d is None:
        raise RuntimeError("reference not vendored: run tools/vendor_reference.sh in the build container")
    try:
        import wandb  # noqa: F401
    except Exception:
        sys.modules["wandb"] = types.ModuleType("wandb")
    import torch
    prec = torch.get_float32_matmul_precision()
    sys.path.insert(0, d)
    try:
        for name in ("model", "layers", "utility_layers", "training_utilities"):
            sys.modules.pop(name, None)
        import model as ref_model
    finally:
        sys.path.remove(d)
        torch.set_float32_matmul_precision(prec)      # model.py:9 sets 'high' as an import side effect
    _mod = ref_model
    return _mod


def build_model(cfg: dict, state_dict=None):
    """`MainModel.from_dict(**cfg)` (utility_layers.py:163-167) in eval mode, optionally with a state_dict
    loaded strictly."""
    m = load().MainModel.from_dict(**cfg)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()
