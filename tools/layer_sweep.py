"""BASELINE.json configs[4]: isolated layer sweep at embed 768 -- one ConvMixer (DW 7x7 + pointwise + MLP)
and one EncoderLayer with 5 register tokens on a 14x14 grid, batch 1..2048, bf16.  Times the engine's
kernel sequence for the layer (token-major buffers, CUDA events around ~25 ms bursts of back-to-back calls after a warm-up burst, median of 5) and prints a
markdown table."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402
import sdpnet_oracle as O  # noqa: E402
from sdpnet_b200.engine import Buffers, Packer, run_encoder, run_mixer  # noqa: E402

C, h, G, R, k = 768, 8, 14, 5, 7
T, S = G * G, G * G + R
cfg = dict(embedding_dim=C, n_head=h, num_blocks=1, conv_kernel_size=k, patch_size=16, conv_block_num=1,
           max_image_size=[16, 16], head_output_from_register=True)
sd = O.synth_state_dict(cfg, seed=0, stress=True)
pk = Packer("cuda", "bf16", C, h, "gelu")
wm = pk._pack_mixer(sd, "blocks.0.conv_blocks.0.")
we = pk._pack_encoder(sd, "blocks.0.t_block.")
mix_flops = 2 * T * C * k * k + 2 * S * C * C + 16 * S * C * C
enc_flops = 6 * S * C * C + 4 * S * S * C + 2 * S * C * C + 16 * S * C * C


def med_ms(fn, reps=5, target_ms=25.0):
    """Median over `reps` of the time per call inside a burst of back-to-back calls about `target_ms` long, after a
    warm-up burst of the same length: steady state.  (A layer inside a network never starts on an idle GPU; single calls
    bracketed by synchronisations measure clock and power-state ramps instead -- up to 1.5x slower at batch 512.)"""
    def burst(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    burst(2)
    inner = max(3, min(400, int(target_ms / max(burst(3), 1e-3))))
    burst(inner)
    ts = sorted(burst(inner) for _ in range(reps))
    return ts[len(ts) // 2]


print("| batch | ConvMixer ms | img/s | TFLOP/s | EncoderLayer ms | img/s | TFLOP/s |")
print("|---:|---:|---:|---:|---:|---:|---:|")
only = [int(v) for v in os.environ.get("SWEEP_BATCHES", "").split(",") if v]     # e.g. SWEEP_BATCHES=512,1024
B = 1
while B <= 2048:
    if only and B not in only:
        B *= 2
        continue
    bufs = Buffers(pk, B, T, R)
    bufs.act.copy_(torch.randn(B, S, C, device="cuda"))
    tm = med_ms(lambda: run_mixer(pk, wm, bufs, G, G))
    te = med_ms(lambda: run_encoder(pk, we, bufs))
    print(f"| {B} | {tm:.3f} | {B / tm * 1e3:.0f} | {B * mix_flops / tm / 1e9:.1f} | {te:.3f} | {B / te * 1e3:.0f} | "
          f"{B * enc_flops / te / 1e9:.1f} |", flush=True)
    del bufs
    B *= 2
