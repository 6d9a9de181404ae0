"""The widened path end to end (GPU box): decoded uint8 images on the host -> H2D -> val_transforms kernels (bf16 out)
-> XL forward (one sdp_forward call) -> on-device loss / top-1 accumulation; one host sync at the end.  Every step of a
batch runs on the GPU; the H2D copy of batch i+1 overlaps the compute of batch i.  Prints one JSON line.
python tools/pipeline_probe.py [batch] [batches]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpnet_b200 as sdp  # noqa: E402
from bench import CONFIGS, NUM_REGISTERS  # noqa: E402

SIZES = [(375, 500), (500, 375), (333, 500), (500, 333), (480, 640), (360, 480), (500, 500), (768, 1024)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cfg, _ = CONFIGS["XL"]
    torch.manual_seed(0)
    model = sdp.MainModel.from_dict(**cfg).eval().to("cuda")
    gi = torch.Generator().manual_seed(5)
    base = [torch.randint(0, 256, (h, w, 3), generator=gi, dtype=torch.uint8).numpy() for (h, w) in SIZES]
    imgs = [base[i % len(base)] for i in range(B)]
    labels = torch.randint(0, cfg["output_classes"], (B,), generator=torch.Generator().manual_seed(0)).cuda()
    tfs = [sdp.val_transforms(out_dtype=torch.bfloat16) for _ in range(2)]       # two pinned staging buffers
    packed = [tf.pack(imgs) for tf in tfs]                                        # host packing is the loader's job
    meter = sdp.evaluate.EvalMeter("cuda")
    copy_stream, main_stream = torch.cuda.Stream(), torch.cuda.current_stream()

    def upload(i):
        host, desc = packed[i % 2]
        with torch.cuda.stream(copy_stream):
            px = host.to("cuda", non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return px, desc, ev

    def compute(i, px, desc, ev):
        main_stream.wait_event(ev)
        px.record_stream(main_stream)
        tf = tfs[i % 2]
        need = sdp.ops.val_preprocess_workspace_bytes(desc, B, tf.image_size, tf.crop_size)
        if tf._ws is None or tf._ws.numel() < need:
            tf._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        x = torch.empty(B, 3, *tf.crop_size, dtype=torch.bfloat16, device="cuda")
        sdp.ops.val_preprocess(px, desc, B, tf.image_size, tf.crop_size, tf.mean, tf.std, tf._ws, x)
        meter.update(model(x, NUM_REGISTERS), labels)

    def run(n):
        nxt = upload(0)
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = upload(i + 1)
            compute(i, *cur)
        return meter.result()

    run(2)
    meter.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = run(nb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "raw uint8 images -> val_transforms -> SdP-Net XL bf16 forward -> on-device metrics, images/sec",
                      "value": B * nb / dt, "unit": "images/s", "batch": B, "batches": nb, "ms_per_batch": 1e3 * dt / nb,
                      "h2d_bytes_per_batch": int(packed[0][0].numel()), "metrics": res}))


if __name__ == "__main__":
    main()
