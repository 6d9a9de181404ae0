"""CPU tests of the validation-preprocessing row: the oracle (oracle/preprocess_oracle.py)
against golden outputs of the REAL torchvision/Pillow pipeline of hf_dataset_generator.py:27-41 (bit-exact), and the
host side of `sdpnet_b200.val_transforms` (packing, planning, argument errors).  No compute call needs a GPU here."""
import glob
import os

import numpy as np
import pytest
import torch

import preprocess_oracle as P
from conftest import GOLDEN

CASES = sorted(os.path.basename(f)[len("preprocess_"):-4] for f in glob.glob(os.path.join(GOLDEN, "preprocess_*.npz"))
               if not f.endswith("preprocess_lut.npz"))


@pytest.fixture(scope="module")
def sdp():
    import sdpnet_b200 as m
    return m


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"preprocess_{name}.npz"))
    img = P.synth_image(int(z["H"]), int(z["W"]), int(z["seed"]))
    return img, tuple(int(v) for v in z["resize"]), tuple(int(v) for v in z["crop"]), z


def test_fixture_set_is_complete():
    assert len(CASES) == 10 and {"imagenet_like", "same_both", "tiny_source", "big_down"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_pipeline_bit_for_bit(name):
    img, rs, cs, z = load_case(name)
    u8 = P.resize_center_crop_u8(img, rs, cs)                        # [h, w, 3]
    assert np.array_equal(u8.transpose(2, 0, 1), z["u8"])            # Pillow resize + crop, every byte
    out = P.val_preprocess(img, rs, cs)
    assert out.dtype == np.float32 and out.shape == (3,) + cs
    ref = z["out"]
    got = out if ref.shape == out.shape else out[:, :8, :8]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))  # float32 bit patterns
    assert float(out.astype(np.float64).sum()) == float(z["out_sum"])


def test_normalisation_table_matches_torchvision():
    lut = np.load(os.path.join(GOLDEN, "preprocess_lut.npz"))["lut"]
    assert np.array_equal(P.normalize_lut().view(np.uint32), lut.view(np.uint32))


def test_coefficients_sum_to_one_and_identity_when_sizes_match():
    k, b, kk = P.precompute_coeffs(40, 40)
    assert k == 5 and all(int(kk[i].sum()) == 1 << 22 for i in range(40))
    for i in range(40):                                              # scale 1: the tap on the pixel itself carries it all
        assert kk[i, i - b[i, 0]] == 1 << 22
    k, b, kk = P.precompute_coeffs(611, 40)
    assert k == int(np.ceil(2 * 611 / 40)) * 2 + 1 and abs(int(kk[7].sum()) - (1 << 22)) <= k


def test_host_plan_agrees_with_oracle(sdp):
    from sdpnet_b200 import _lib as L
    sizes = [(375, 500), (500, 333), (37, 53), (2000, 1500), (320, 320), (3, 2)]
    desc = (L.ImageDesc * len(sizes))()
    off = 0
    for i, (h, w) in enumerate(sizes):
        desc[i].offset, desc[i].height, desc[i].width = off, h, w
        off += (3 * h * w + 15) // 16 * 16
    rs, cs = (320, 320), (224, 224)
    kmax, rows = P.plan([s[0] for s in sizes], [s[1] for s in sizes], rs, cs)
    up = lambda v: (v + 255) // 256 * 256
    want = up(16 * len(sizes)) + up(len(sizes) * (cs[0] + cs[1]) * (kmax + 2) * 4) + up(768 * 4) + up(len(sizes) * rows * cs[1] * 3)
    assert sdp.ops.val_preprocess_workspace_bytes(desc, len(sizes), rs, cs) == want


def test_pack_layout_and_argument_errors(sdp):
    t = sdp.val_transforms()                                         # reference defaults
    assert t.image_size == (320, 320) and t.crop_size == (224, 224) and t.mean == [0.485, 0.456, 0.406]
    a, b = P.synth_image(5, 7, 1), P.synth_image(9, 4, 2)
    host, desc = t.pack([a, b])
    assert [(d.offset, d.height, d.width) for d in desc] == [(0, 5, 7), (112, 9, 4)]
    assert np.array_equal(host.numpy()[:105], a.reshape(-1)) and np.array_equal(host.numpy()[112:220], b.reshape(-1))
    with pytest.raises(ValueError):
        sdp.val_transforms((20, 20), (24, 24))                       # crop larger than the resize
    with pytest.raises(TypeError):
        sdp.val_transforms(320, 224)                                 # smaller-edge form is not val_transforms
    with pytest.raises(ValueError):
        sdp.val_transforms(std=[0.2, 0.0, 0.2])
    with pytest.raises(TypeError):
        t([np.zeros((4, 4), np.uint8)])                              # not RGB
    with pytest.raises(TypeError):
        t([np.zeros((4, 4, 3), np.float32)])


def test_pil_inputs_are_converted_like_transforms_rgb(sdp):
    from PIL import Image
    from sdpnet_b200.preprocess import _as_rgb_u8
    g = Image.fromarray(np.arange(12, dtype=np.uint8).reshape(3, 4), "L")
    assert np.array_equal(_as_rgb_u8(g), np.asarray(g.convert("RGB")))
    rgba = Image.fromarray(P.synth_image(4, 4, 3)).convert("RGBA")
    assert np.array_equal(_as_rgb_u8(rgba), np.asarray(rgba.convert("RGB")))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(sdp):
    with pytest.raises(Exception):
        sdp.val_transforms(device="cpu")([P.synth_image(8, 8, 0)])
