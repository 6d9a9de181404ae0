"""Validation-preprocessing probe (GPU box): times sdp_val_preprocess on an ImageNet-shaped batch (device-resident
packed pixels, CUDA events) and end to end from host arrays, next to the reference pipeline itself (torchvision v2 +
Pillow, hf_dataset_generator.py:27-41) on one host core.  Prints one JSON line.
python tools/preprocess_probe.py [batch] [--bf16]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import preprocess_oracle as P  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402

SIZES = [(375, 500), (500, 375), (333, 500), (500, 333), (480, 640), (360, 480), (500, 500), (768, 1024)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1024
    dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
    base = [P.synth_image(h, w, i) for i, (h, w) in enumerate(SIZES)]
    imgs = [base[i % len(base)] for i in range(B)]
    t = sdp.val_transforms(out_dtype=dt)
    host, desc = t.pack(imgs)
    pixels = host.cuda()
    need = sdp.ops.val_preprocess_workspace_bytes(desc, B, t.image_size, t.crop_size)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, 3, 224, 224, dtype=dt, device="cuda")
    run = lambda: sdp.ops.val_preprocess(pixels, desc, B, t.image_size, t.crop_size, t.mean, t.std, ws, out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 20
    ev[0].record()
    for _ in range(n):
        run()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    # end to end: host arrays -> pinned pack -> H2D -> kernels (the call a loader makes)
    t0 = time.perf_counter()
    for _ in range(3):
        t(imgs)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    # algorithmic bytes: the source window each image's crop needs, read once, plus the output written once
    in_bytes = 0
    for (h, w) in [im.shape[:2] for im in imgs]:
        _, bv, _ = P.precompute_coeffs(h, 320)
        _, bh, _ = P.precompute_coeffs(w, 320)
        rows = bv[271].sum() - bv[48, 0]
        cols = bh[271].sum() - bh[48, 0]
        in_bytes += int(rows) * int(cols) * 3
    out_bytes = out.numel() * out.element_size()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = float(peaks.get("hbm_gbs", peaks.get("hbm_gbps", 6456.2))) if isinstance(peaks, dict) else 6456.2
    # the reference itself on one host core (bounded sample)
    cpu = None
    try:
        import torchvision.transforms.v2 as T
        from PIL import Image
        ref = T.Compose([T.RGB(), T.Resize((320, 320), interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop((224, 224)),
                         T.ToImage(), T.ToDtype(torch.float32, scale=True), T.Normalize(t.mean, t.std)])
        pil = [Image.fromarray(a, "RGB") for a in base]
        torch.set_num_threads(1)
        ref(pil[0])
        t0 = time.perf_counter()
        k = 0
        while time.perf_counter() - t0 < 5.0:
            r = ref(pil[k % len(pil)])
            k += 1
        cpu = k / (time.perf_counter() - t0)
        same = bool(torch.equal(r, t(base[(k - 1) % len(pil)]).float().cpu())) if dt == torch.float32 else None
    except Exception as e:  # noqa: BLE001
        cpu, same = None, repr(e)
    print(json.dumps({"metric": "val preprocessing images/sec", "batch": B, "out_dtype": str(dt), "ms_per_batch": ms,
                      "value": B / ms * 1e3, "e2e_from_host_arrays_images_per_s": B / e2e_ms * 1e3,
                      "h2d_bytes": int(host.numel()), "algorithmic_bytes": in_bytes + out_bytes,
                      "roofline": {"bound": "hbm", "achieved": (in_bytes + out_bytes) / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                   "frac": (in_bytes + out_bytes) / ms / 1e6 / hbm},
                      "cpu_reference": {"value": cpu, "unit": "images/s", "cores": 1, "kind": "reference",
                                        "sample": "torchvision v2 + Pillow val_transforms on PIL images, 5 s, one thread",
                                        "matches_gpu_bit_for_bit": same}}))


if __name__ == "__main__":
    main()
