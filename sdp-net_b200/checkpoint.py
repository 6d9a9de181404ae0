"""Checkpoint loaders for the reference's on-disk formats -- host logic only.

Formats (all `torch.save`d dicts):
* Trainer snapshot  {"model_state_dict", "model_config", "optimizer_state", "scheduler_state", "epoch"}
  (reference training_tools.py:203-226); keys carry a "module." prefix when saved from DDP.
* EMA weights       a bare state_dict (training_tools.py:282-302, `ema_model.pt`).
* save_model file   {"state_dict", "config"} (utility_layers.py:185-198, read back by from_pretrained :169-183).
Prefix stripping mirrors `preprocess_weights` (model_test.py:21-26) and also handles the "_orig_mod."
prefix a `torch.compile`d model adds (which nothing upstream strips).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

PREFIXES = ("module.", "_orig_mod.")


def strip_prefixes(sd: Dict[str, torch.Tensor], prefixes=PREFIXES) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        changed = True
        while changed:
            changed = False
            for p in prefixes:
                if k.startswith(p):
                    k, changed = k[len(p):], True
        out[k] = v.detach().to("cpu") if isinstance(v, torch.Tensor) else v
    return out


def read_checkpoint(obj_or_path, map_location="cpu") -> Tuple[Optional[dict], Dict[str, torch.Tensor], dict]:
    """-> (model_config or None, state_dict with prefixes stripped, extras such as epoch)."""
    blob = torch.load(obj_or_path, map_location=map_location) if isinstance(obj_or_path, (str, bytes)) or \
        hasattr(obj_or_path, "read") else obj_or_path
    if not isinstance(blob, dict):
        raise TypeError("checkpoint is not a dict")
    if "model_state_dict" in blob:                                   # Trainer snapshot
        extras = {k: blob[k] for k in ("epoch",) if k in blob}
        return dict(blob.get("model_config") or {}) or None, strip_prefixes(blob["model_state_dict"]), extras
    if "state_dict" in blob and "config" in blob:                    # SdPModel.save_model
        return dict(blob["config"]) or None, strip_prefixes(blob["state_dict"]), {}
    if all(isinstance(v, torch.Tensor) for v in blob.values()):      # EMA file / bare state_dict
        return None, strip_prefixes(blob), {}
    raise ValueError("unrecognised checkpoint layout: keys " + ", ".join(sorted(map(str, blob))[:8]))


def load_model(path, ema_path=None, device="cuda", precision="bf16", config: Optional[dict] = None):
    """model_test.return_model (model_test.py:28-42) on the engine: returns `model` or `(model, ema_model)`."""
    from .model import MainModel

    cfg, sd, _ = read_checkpoint(path)
    cfg = cfg or config
    if cfg is None:
        raise ValueError("checkpoint carries no model_config; pass config=")

    def build(state):
        m = MainModel.from_dict(**cfg)
        m.load_state_dict(state, strict=True)
        m = m.eval()
        if device is not None:
            m = m.to(device)
        return m.set_precision(precision)

    model = build(sd)
    if ema_path is None:
        return model
    _, ema_sd, _ = read_checkpoint(ema_path)
    return model, build(ema_sd)
