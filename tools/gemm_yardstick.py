"""Yardstick (GPU box): the XL GEMM shapes through sdp_gemm and through cuBLAS (torch.matmul / F.linear), each run
back to back for ~1.5 s so both see the same power-capped clocks.  cuBLAS is NOT used by the product; this only says
how far the hand-written kernel is from the library at the same shape.   python tools/gemm_yardstick.py [B]"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp  # noqa: E402


class Sampler(threading.Thread):
    """nvidia-smi SM clock / power polled from a side thread, so the timed loop never waits for it."""

    def __init__(self):
        super().__init__(daemon=True)
        self.stop, self.rows = threading.Event(), []

    def run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(o[0]), float(o[1])))
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.05)

    def summary(self):
        self.stop.set()
        self.join()
        r = self.rows[len(self.rows) // 3:] or self.rows      # drop the ramp
        if not r:
            return "?"
        return f"{statistics.median(x[0] for x in r):.0f} MHz {statistics.median(x[1] for x in r):.0f} W"


def sustained(fn, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    smp = Sampler()
    smp.start()
    n, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        if n % 100 == 0:                                       # bound the launch queue; the GPU idles ~10 us per 100 calls
            torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, smp.summary()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    M = B * 261
    g = torch.Generator(device="cuda").manual_seed(0)
    print(f"M = {M}")
    for (N, K, act, res) in [(768, 768, "none", True), (768, 768, "gelu", True), (3072, 768, "gelu", False),
                             (768, 3072, "none", True), (2304, 768, "none", False)]:
        A = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        resid = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
        fl = 2.0 * M * N * K
        ms_sdp, c1 = sustained(lambda: sdp.ops.gemm(A, W, out, bias=bias, act=act, residual=resid))
        ms_plain, c2 = sustained(lambda: sdp.ops.gemm(A, W, out))
        ms_cb, c3 = sustained(lambda: torch.matmul(A, W.t(), out=out))
        bb = bias.bfloat16()
        ms_lin, c4 = sustained(lambda: F.linear(A, W, bb))
        print(f"N{N} K{K} act={act} res={res}: sdp(epi) {ms_sdp:.3f} ms {fl / ms_sdp / 1e9:7.1f} TF/s [{c1}] | sdp(plain) {ms_plain:.3f} ms "
              f"{fl / ms_plain / 1e9:7.1f} [{c2}] | cuBLAS matmul {ms_cb:.3f} ms {fl / ms_cb / 1e9:7.1f} [{c3}] | "
              f"cuBLAS linear+bias {ms_lin:.3f} ms {fl / ms_lin / 1e9:7.1f} [{c4}]", flush=True)
        del A, W, out, resid
    return 0


if __name__ == "__main__":
    sys.exit(main())
