"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's validation preprocessing,
`val_transforms` in /root/reference/hf_dataset_generator.py:27-41:

    RGB() -> Resize((320, 320), BICUBIC) -> CenterCrop((224, 224)) -> ToImage() -> ToDtype(float32, scale=True)
          -> Normalize(mean, std)

applied to a PIL image, i.e. the resize is Pillow's `Image.resize(..., resample=BICUBIC)` on 8-bit RGB data
(torchvision `_resize_image_pil`, transforms/v2/functional/_geometry.py), NOT torch's interpolate.  Pillow is a
third-party dependency of the reference (unpinned: the reference has no requirements file; the build container has
Pillow 12.2.0 and torchvision 0.26.0).  Its published algorithm (src/libImaging/Resample.c: `precompute_coeffs`,
`normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`, `ImagingResample`) is
restated here in numpy:

  * per axis and output index: centre = (i + 0.5) * scale, support = 2 * max(scale, 1), taps
    [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the image, Keys bicubic weights (a = -0.5)
    evaluated in double precision, normalised to sum 1, then rounded to 22-bit fixed point;
  * horizontal pass over the source rows the vertical pass needs, result clipped to uint8; then the vertical pass,
    again clipped to uint8 (two roundings, exactly as Pillow);
  * centre crop anchor int(round((resized - crop) / 2.0)) (torchvision `_center_crop_compute_crop_anchor`);
  * float32(v) * float32(1/255), minus float32 mean, divided by float32 std (torchvision `to_dtype_image`,
    `normalize_image`).

Pinned by `oracle/make_golden_preprocess.py`, which runs the real torchvision/Pillow pipeline in the build container
and commits its outputs to tests/golden/preprocess_*.npz; tests/test_preprocess_cpu.py holds this file to them
bit for bit.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)     # hf_dataset_generator.py:30
IMAGENET_STD = (0.229, 0.224, 0.225)      # hf_dataset_generator.py:31


def synth_image(h: int, w: int, seed: int) -> np.ndarray:
    """Deterministic uint8 RGB test image: smooth gradients + hashed noise + hard edges (integer arithmetic only)."""
    y, x, c = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), np.arange(3, dtype=np.uint64),
                          indexing="ij")
    i = (y * np.uint64(w) + x) * np.uint64(3) + c
    hsh = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    hsh = (hsh ^ (hsh >> np.uint64(15))) * np.uint64(2246822519) & np.uint64(0xFFFFFFFF)
    noise = (hsh >> np.uint64(13)) & np.uint64(0xFF)
    grad = (x * np.uint64(255) // np.uint64(max(w - 1, 1)) + y * np.uint64(200) // np.uint64(max(h - 1, 1)) + c * np.uint64(40)) & np.uint64(0xFF)
    edges = np.where(((x // np.uint64(7)) + (y // np.uint64(5))) % np.uint64(2) == 0, np.uint64(255), np.uint64(0))
    sel = (hsh >> np.uint64(24)) % np.uint64(3)
    img = np.where(sel == 0, noise, np.where(sel == 1, grad, edges))
    return img.astype(np.uint8)


def bicubic_filter(x: float) -> float:
    """Keys cubic convolution kernel with a = -0.5 (Resample.c `bicubic_filter`), same operation order."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the full box (0, in_size).
    Returns (ksize, bounds[out_size, 2] = (first tap, tap count), kk[out_size, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, first: int, last: int) -> np.ndarray:
    """out[i] = clip8(2^21 + sum_k img[bounds[i,0] + k] * kk[i, k]) for i in [first, last); img is [n, ...] uint8."""
    out = np.empty((last - first,) + img.shape[1:], np.uint8)
    for i in range(first, last):
        lo, cnt = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(cnt):
            acc += img[lo + k].astype(np.int64) * int(kk[i, k])
        out[i - first] = _clip8(acc)
    return out


def crop_anchor(resized: int, crop: int) -> int:
    return int(round((resized - crop) / 2.0))


def resize_center_crop_u8(img: np.ndarray, resize=(320, 320), crop=(224, 224)) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [crop_h, crop_w, 3]: Pillow bicubic resize to `resize` then the centre crop.
    Only the crop window is computed (every output pixel depends on its own taps only, so the values equal the
    window of the full resize)."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    H, W = img.shape[:2]
    (rh, rw), (ch, cw) = resize, crop
    assert ch <= rh and cw <= rw, "crop larger than the resized image (torchvision would pad): not supported"
    top, left = crop_anchor(rh, ch), crop_anchor(rw, cw)
    _, bh, kh = precompute_coeffs(W, rw)
    _, bv, kv = precompute_coeffs(H, rh)
    if (rh, rw) == (H, W):            # Image.resize returns a copy
        return img[top:top + ch, left:left + cw].copy()
    # horizontal pass on the source rows the cropped vertical pass needs (Resample.c computes rows
    # [bounds_vert[0], last) for the whole output; the rows used by the crop window are a subset with equal values)
    y0 = int(bv[top, 0])
    y1 = int(bv[top + ch - 1, 0] + bv[top + ch - 1, 1])
    rows = img[y0:y1]
    if W != rw:                       # need_horizontal
        tmp = _resample_axis0(np.ascontiguousarray(rows.transpose(1, 0, 2)), bh, kh, left, left + cw).transpose(1, 0, 2)
    else:
        tmp = rows[:, left:left + cw]
    if H != rh:                       # need_vertical
        bv2 = bv.copy()
        bv2[:, 0] -= y0
        out = _resample_axis0(np.ascontiguousarray(tmp), bv2, kv, top, top + ch)
    else:
        out = tmp[top - y0:top - y0 + ch]
    return np.ascontiguousarray(out)


def normalize_lut(mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """[3, 256] float32: (float32(v) * float32(1/255) - mean_c) / std_c with every step rounded to float32."""
    v = np.arange(256, dtype=np.float32) * np.float32(1.0 / 255)
    m = np.asarray(mean, np.float32)[:, None]
    s = np.asarray(std, np.float32)[:, None]
    return ((v[None, :] - m) / s).astype(np.float32)


def val_preprocess(img: np.ndarray, resize=(320, 320), crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """uint8 RGB [H, W, 3] -> float32 [3, crop_h, crop_w], the tensor `val_transforms()(pil_image)` returns."""
    u8 = resize_center_crop_u8(img, resize, crop)
    lut = normalize_lut(mean, std)
    out = np.empty((3,) + u8.shape[:2], np.float32)
    for c in range(3):
        out[c] = lut[c][u8[:, :, c]]
    return out


def plan(heights, widths, resize=(320, 320), crop=(224, 224)):
    """(kmax, temp_rows): the largest tap count and the most source rows any image's crop window needs — what the
    host sizes the coefficient and intermediate buffers with."""
    kmax, rows = 1, 1
    top = crop_anchor(resize[0], crop[0])
    for H, W in zip(heights, widths):
        kw, _, _ = precompute_coeffs(W, resize[1])
        kh, bv, _ = precompute_coeffs(H, resize[0])
        kmax = max(kmax, kw, kh)
        rows = max(rows, int(bv[top + crop[0] - 1].sum() - bv[top, 0]))
    return kmax, rows
