"""im2col of the patch embedding in isolation (GPU box): ms per launch, GB/s of (input read + patch matrix written), exactness
against torch unfold.  python tools/im2col_probe.py"""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpnet_b200 as sdp
for (B, H, p, dt) in [(1024, 224, 14, torch.float32), (1024, 224, 14, torch.bfloat16), (512, 224, 16, torch.float32)]:
    x = torch.randn(B, 3, H, H, device='cuda').to(dt)
    G = H // p
    ld = (3 * p * p + 7) // 8 * 8
    A = torch.empty(B * G * G, ld, device='cuda', dtype=torch.bfloat16)
    f = lambda: sdp.ops.im2col_patches(x, A, p)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = (x.numel() * x.element_size() + A.numel() * 2) / 1e9
    ref = torch.nn.functional.unfold(x.float(), p, stride=p).transpose(1, 2).reshape(B * G * G, 3 * p * p).bfloat16()
    ok = torch.equal(A[:, :3 * p * p], ref) and bool((A[:, 3 * p * p:] == 0).all())
    print(f"B{B} {H}^2 p{p} {dt}: {ms:.3f} ms {gb / ms * 1e3:.0f} GB/s exact={ok}")
