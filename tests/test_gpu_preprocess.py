"""GPU parity of the validation-preprocessing kernels (sdp_val_preprocess, through the C-ABI) — bit-exact against
the golden outputs of the real torchvision/Pillow pipeline and against the oracle on batches of ragged sizes."""
import glob
import os

import numpy as np
import pytest
import torch

import preprocess_oracle as P
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(f)[len("preprocess_"):-4] for f in glob.glob(os.path.join(GOLDEN, "preprocess_*.npz"))
               if not f.endswith("preprocess_lut.npz"))


@pytest.fixture(scope="module")
def sdp():
    import sdpnet_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", CASES)
def test_golden_fixture_bit_exact(sdp, name, path):
    z = np.load(os.path.join(GOLDEN, f"preprocess_{name}.npz"))
    img = P.synth_image(int(z["H"]), int(z["W"]), int(z["seed"]))
    rs, cs = tuple(int(v) for v in z["resize"]), tuple(int(v) for v in z["crop"])
    out = sdp.val_transforms(rs, cs)(img).cpu().numpy()              # single image -> [3, h, w]
    lut = np.load(os.path.join(GOLDEN, "preprocess_lut.npz"))["lut"]
    want = np.stack([lut[c][z["u8"][c]] for c in range(3)])          # real pipeline's bytes through its own table
    assert out.shape == want.shape and np.array_equal(bits(out), bits(want))
    ref = z["out"]
    got = out if ref.shape == out.shape else out[:, :8, :8]
    assert np.array_equal(bits(got), bits(ref))
    assert float(out.astype(np.float64).sum()) == float(z["out_sum"])


def test_ragged_batch_matches_oracle(sdp, path):
    sizes = [(37, 53), (150, 97), (64, 40), (40, 91), (40, 40), (1, 1), (2, 300), (300, 2), (129, 131), (611, 807), (33, 40)]
    imgs = [P.synth_image(h, w, 100 + i) for i, (h, w) in enumerate(sizes)]
    rs, cs = (40, 40), (28, 28)
    out = sdp.val_transforms(rs, cs)(imgs).cpu().numpy()
    assert out.shape == (len(sizes), 3, 28, 28)
    for i, im in enumerate(imgs):
        assert np.array_equal(bits(out[i]), bits(P.val_preprocess(im, rs, cs))), sizes[i]


@pytest.fixture(params=["fused", "general"])
def path(request, monkeypatch, sdp):
    """Both kernel paths: the fused band kernel (default) and the three-kernel path with a global intermediate (the
    `path` argument of sdp_val_preprocess, here through ValTransforms.general_path)."""
    monkeypatch.setattr(sdp.preprocess.ValTransforms, "general_path", request.param == "general")
    return request.param


def test_odd_crop_width_and_custom_statistics(sdp, path):
    rs, cs, mean, std = (48, 36), (31, 21), [0.1, 0.5, 0.9], [0.5, 0.25, 2.0]      # 21 * 3 bytes per row: byte path
    imgs = [P.synth_image(75, 61, 5), P.synth_image(20, 90, 6)]
    out = sdp.val_transforms(rs, cs, mean, std)(imgs).cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(bits(out[i]), bits(P.val_preprocess(im, rs, cs, mean, std)))


def test_imagenet_sized_batch(sdp, path):
    sizes = [(375, 500), (500, 333), (333, 500), (480, 640), (1200, 1600), (224, 224), (320, 320), (64, 64)] * 2
    imgs = [P.synth_image(h, w, 200 + i) for i, (h, w) in enumerate(sizes)]
    t = sdp.val_transforms()
    out = t(imgs)
    assert out.shape == (len(sizes), 3, 224, 224) and out.dtype == torch.float32
    o = out.cpu().numpy()
    for i in (0, 1, 4, 6, 7, 12):                                     # the oracle is slow: a sample of the batch
        assert np.array_equal(bits(o[i]), bits(P.val_preprocess(imgs[i]))), sizes[i]
    # permutation of the batch permutes the outputs, and nothing else (per-image independence)
    perm = [5, 3, 0, 9, 1]
    o2 = t([imgs[j] for j in perm]).cpu().numpy()
    for k, j in enumerate(perm):
        assert np.array_equal(bits(o2[k]), bits(o[j]))
    # an image that already has the resize size only goes through crop + table
    lut = P.normalize_lut()
    direct = np.stack([lut[c][imgs[6][48:272, 48:272, c]] for c in range(3)])
    assert np.array_equal(bits(o[6]), bits(direct))


def test_huge_downscale_falls_back_to_the_general_path(sdp):
    # 9000 columns -> 40: 1803 taps per output column; the tap table alone exceeds shared memory, so no band fits
    im = P.synth_image(12, 9000, 3)
    out = sdp.val_transforms((40, 40), (28, 28))([im, P.synth_image(30, 30, 4)]).cpu().numpy()
    assert np.array_equal(bits(out[0]), bits(P.val_preprocess(im, (40, 40), (28, 28))))


def test_bf16_output_is_the_rounded_float_output(sdp, path):
    imgs = [P.synth_image(90, 120, 1), P.synth_image(50, 45, 2)]
    f = sdp.val_transforms((64, 64), (48, 48))(imgs)
    h = sdp.val_transforms((64, 64), (48, 48), out_dtype=torch.bfloat16)(imgs)
    assert h.dtype == torch.bfloat16 and torch.equal(h, f.to(torch.bfloat16))


def test_constant_image_and_feeds_the_model(sdp):
    const = np.full((77, 91, 3), 200, np.uint8)
    out = sdp.val_transforms((40, 40), (28, 28))(const).cpu().numpy()
    lut = P.normalize_lut()
    for c in range(3):
        assert np.all(out[c] == lut[c][200])                          # weights sum to one: constants survive both passes
    import sdpnet_oracle as O
    cfg = dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4, output_classes=10, max_image_size=[8, 8])
    model = sdp.MainModel.from_dict(**cfg).eval().to("cuda")
    model.load_state_dict(O.synth_state_dict(cfg, seed=1))
    x = sdp.val_transforms((40, 40), (32, 32))([P.synth_image(60, 50, 3), P.synth_image(45, 80, 4)])
    logits = model(x, 3)
    assert logits.shape == (2, 10) and bool(torch.isfinite(logits.float()).all())


def test_unpadded_pixel_buffer_ending_mid_word(sdp, path):
    """C-ABI callers need not pad: the last image may end at any byte (the kernels never read past pixels_bytes)."""
    from sdpnet_b200 import _lib as L
    a, b = P.synth_image(5, 7, 11), P.synth_image(9, 5, 12)          # 105 + 135 bytes, second image at an odd offset
    px = torch.from_numpy(np.concatenate([a.reshape(-1), b.reshape(-1)])).cuda()
    desc = (L.ImageDesc * 2)()
    desc[0].offset, desc[0].height, desc[0].width = 0, 5, 7
    desc[1].offset, desc[1].height, desc[1].width = 105, 9, 5
    rs, cs = (40, 40), (28, 28)
    ws = torch.empty(sdp.ops.val_preprocess_workspace_bytes(desc, 2, rs, cs), dtype=torch.uint8, device="cuda")
    out = torch.empty(2, 3, 28, 28, device="cuda")
    sdp.ops.val_preprocess(px, desc, 2, rs, cs, list(P.IMAGENET_MEAN), list(P.IMAGENET_STD), ws, out)
    o = out.cpu().numpy()
    assert np.array_equal(bits(o[0]), bits(P.val_preprocess(a, rs, cs)))
    assert np.array_equal(bits(o[1]), bits(P.val_preprocess(b, rs, cs)))


def test_bad_descriptors_raise(sdp):
    from sdpnet_b200 import _lib as L
    desc = (L.ImageDesc * 1)()
    desc[0].offset, desc[0].height, desc[0].width = 0, 50, 50
    px = torch.zeros(100, dtype=torch.uint8, device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty(1, 3, 28, 28, device="cuda")
    with pytest.raises(L.SdpNetLibraryError):
        sdp.ops.val_preprocess(px, desc, 1, (40, 40), (28, 28), [0.5] * 3, [0.5] * 3, ws, out)   # pixels too short
    px = torch.zeros(7500, dtype=torch.uint8, device="cuda")
    with pytest.raises(L.SdpNetLibraryError):
        sdp.ops.val_preprocess(px, desc, 1, (40, 40), (28, 28), [0.5] * 3, [0.5] * 3, ws[:64], out)  # workspace too small
