"""Validation preprocessing on the GPU, the step right in front of the forward.

Mirrors `val_transforms(image_size, crop_size, mean, std)` of the reference (hf_dataset_generator.py:27-41): same
name, arguments and defaults, and the returned object is called the same way — on one PIL image it returns the
`[3, crop_h, crop_w]` float32 tensor the reference transform returns (bit-identical), on a list of images the stacked
batch `[B, 3, crop_h, crop_w]`, which is what the reference's DataLoader collates.  The work the reference does per
sample on CPU workers (Pillow bicubic resize, crop, scale, normalise) happens in three CUDA kernels over the whole
batch (`sdp_val_preprocess`); the host only decodes / converts to 8-bit RGB (`transforms.RGB()`), packs the raw pixels
into one pinned buffer and copies it to the device.  `out_dtype=torch.bfloat16` emits the engine's input dtype directly.
No CPU fallback: without the CUDA library or a GPU this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops

IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]


def _pair(v, what: str):
    if isinstance(v, int):
        raise TypeError(f"{what}: an (h, w) pair is required (the reference passes (320, 320) / (224, 224)); "
                        "the smaller-edge form Resize(int) is not part of val_transforms")
    h, w = (int(x) for x in v)
    if h <= 0 or w <= 0:
        raise ValueError(f"{what} must be positive, got {v}")
    return h, w


def _as_rgb_u8(img) -> np.ndarray:
    """`transforms.RGB()` + raw bytes: PIL image / uint8 [H, W, 3] array or tensor -> contiguous uint8 [H, W, 3]."""
    if hasattr(img, "convert") and hasattr(img, "size"):          # PIL.Image without importing PIL here
        a = np.asarray(img if img.mode == "RGB" else img.convert("RGB"))
    elif isinstance(img, torch.Tensor):
        a = img.detach().cpu().numpy()
    else:
        a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise TypeError(f"val_transforms: expected a PIL image or a uint8 [H, W, 3] RGB array, got {a.dtype} {a.shape}")
    return np.ascontiguousarray(a)


class ValTransforms:
    general_path = False     # True: the three-kernel path of sdp_val_preprocess (same bits; what oversized batches take anyway)

    def __init__(self, image_size=(320, 320), crop_size=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD,
                 out_dtype: torch.dtype = torch.float32, device="cuda"):
        self.image_size = _pair(image_size, "image_size")
        self.crop_size = _pair(crop_size, "crop_size")
        if self.crop_size[0] > self.image_size[0] or self.crop_size[1] > self.image_size[1]:
            raise ValueError("crop_size larger than image_size (the reference would zero-pad): not supported")
        if len(mean) != 3 or len(std) != 3:
            raise ValueError("mean and std need three values (RGB)")
        if any(float(s) == 0.0 for s in std):
            raise ValueError("std evaluated to zero, leading to division by zero.")      # torchvision's message
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("out_dtype must be torch.float32 or torch.bfloat16")
        self.mean = [float(m) for m in mean]
        self.std = [float(s) for s in std]
        self.out_dtype = out_dtype
        self.device = torch.device(device)
        self._ws = None                                            # device workspace, grown on demand
        self._pinned = None                                        # pinned staging buffer, grown on demand
        self._copied = None                                        # event: the last H2D copy out of it has finished

    def __call__(self, images):
        single = not isinstance(images, (list, tuple))
        arrs = [_as_rgb_u8(im) for im in ([images] if single else images)]
        if not arrs:
            raise ValueError("val_transforms: empty batch")
        out = self.run_packed(*self.pack(arrs))
        return out[0] if single else out

    # -- the two halves of __call__, separately usable by a loader that packs on worker threads --
    def pack(self, arrs: Sequence[np.ndarray]):
        """uint8 [H, W, 3] arrays -> (pinned uint8 buffer, descriptor array).  Images start at 16-byte offsets."""
        desc = (L.ImageDesc * len(arrs))()
        off = 0
        for i, a in enumerate(arrs):
            desc[i].offset, desc[i].height, desc[i].width = off, a.shape[0], a.shape[1]
            off = (off + a.size + 15) // 16 * 16
        if self._copied is not None:                                # the previous batch may still be on its way out
            self._copied.synchronize()
        if self._pinned is None or self._pinned.numel() < off:
            self._pinned = torch.empty(max(off, 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        host = self._pinned[:off]
        hv = host.numpy()
        for i, a in enumerate(arrs):
            hv[desc[i].offset:desc[i].offset + a.size] = a.reshape(-1)
        return host, desc

    def run_packed(self, host: torch.Tensor, desc) -> torch.Tensor:
        B = len(desc)
        pixels = host.to(self.device, non_blocking=True)
        if host.is_pinned():
            self._copied = torch.cuda.Event()
            self._copied.record()
        need = ops.val_preprocess_workspace_bytes(desc, B, self.image_size, self.crop_size, self.general_path)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, 3, *self.crop_size, dtype=self.out_dtype, device=self.device)
        return ops.val_preprocess(pixels, desc, B, self.image_size, self.crop_size, self.mean, self.std, self._ws, out,
                                  self.general_path)


def val_transforms(image_size=(320, 320), crop_size=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD,
                   out_dtype: torch.dtype = torch.float32, device="cuda") -> ValTransforms:
    """hf_dataset_generator.py:27-41 `val_transforms`: same positional arguments and defaults."""
    return ValTransforms(image_size, crop_size, mean, std, out_dtype, device)
