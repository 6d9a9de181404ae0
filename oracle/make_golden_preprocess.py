"""Generates tests/golden/preprocess_*.npz by running the REAL validation transform of the reference — the
torchvision v2 Compose of /root/reference/hf_dataset_generator.py:27-41 on PIL images (Pillow does the bicubic
resize) — in the build container.  The reference module itself cannot be imported here (it imports `datasets`,
which is absent), so the Compose is rebuilt verbatim from those lines.  Run once: python oracle/make_golden_preprocess.py
Inputs are procedural (`synth_image`), so the fixtures only hold the expected outputs.
"""
import os
import sys

import numpy as np
import torch
import torchvision.transforms.v2 as transforms
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import preprocess_oracle as P  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


synth_image = P.synth_image


def val_transforms(image_size=(320, 320), crop_size=(224, 224), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    # hf_dataset_generator.py:33-40, verbatim
    return transforms.Compose([
        transforms.RGB(),
        transforms.Resize(image_size, interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.CenterCrop(crop_size),
        transforms.ToImage(),
        transforms.ToDtype(torch.float32, scale=True),
        transforms.Normalize(mean, std),
    ])


# name: (H, W, seed, resize, crop)
CASES = {
    "up_odd": (37, 53, 1, (40, 40), (28, 28)),
    "down_mixed": (150, 97, 2, (40, 40), (28, 28)),
    "same_width": (64, 40, 3, (40, 40), (28, 28)),
    "same_height": (40, 91, 4, (40, 40), (28, 28)),
    "same_both": (40, 40, 5, (40, 40), (28, 28)),
    "nonsquare_cfg": (75, 61, 6, (48, 36), (31, 20)),
    "tiny_source": (3, 2, 7, (40, 40), (28, 28)),
    "big_down": (611, 807, 8, (40, 40), (28, 28)),
    "imagenet_like": (375, 500, 9, (320, 320), (224, 224)),
    "imagenet_tall": (500, 333, 10, (320, 320), (224, 224)),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    worst = 0
    for name, (h, w, seed, rs, cs) in CASES.items():
        img = synth_image(h, w, seed)
        t = val_transforms(rs, cs)(Image.fromarray(img, "RGB"))
        ref = t.numpy()
        # the uint8 image before ToDtype/Normalize, through the same torchvision ops
        u8 = transforms.Compose([transforms.Resize(rs, interpolation=transforms.InterpolationMode.BICUBIC),
                                 transforms.CenterCrop(cs), transforms.ToImage()])(Image.fromarray(img, "RGB")).numpy()
        mine_u8 = P.resize_center_crop_u8(img, rs, cs)
        mine = P.val_preprocess(img, rs, cs)
        d8 = int(np.abs(mine_u8.transpose(2, 0, 1).astype(int) - u8.astype(int)).max())
        same = np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
        worst = max(worst, d8)
        print(f"{name}: {h}x{w} -> {rs} -> {cs}: oracle u8 max diff {d8}, float bit-identical {same}")
        np.savez_compressed(os.path.join(OUT, f"preprocess_{name}.npz"), H=h, W=w, seed=seed, resize=np.array(rs),
                            crop=np.array(cs), u8=u8, out=ref if ref.size <= 4096 else ref[:, :8, :8],
                            out_sum=np.float64(ref.astype(np.float64).sum()))
    lut = np.stack([val_transforms((1, 256), (1, 256))(Image.fromarray(np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, 2), "RGB")).numpy()[c, 0]
                    for c in range(3)])
    print("lut identical:", np.array_equal(lut.view(np.uint32), P.normalize_lut().view(np.uint32)))
    np.savez_compressed(os.path.join(OUT, "preprocess_lut.npz"), lut=lut)
    print("worst u8 diff", worst)


if __name__ == "__main__":
    main()
