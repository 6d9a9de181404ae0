import sys, os
sys.path.insert(0, "/root/repo")
import torch, sdpnet_b200 as sdp
ops = sdp.ops
B, S, C = 1024, 261, 768
M = B * S
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(M, C, device="cuda", generator=g)).bfloat16()
w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.03).bfloat16()
b1 = torch.randn(4 * C, device="cuda", generator=g)
hid = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)
for act in [None, "relu", "gelu", "leaky_relu", "sigmoid", "gelu_tanh"]:
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(x, w1, hid, bias=b1, act=act); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[1:])
    print(f"{str(act):12s} {ts[len(ts)//2]:.3f} ms  {2*M*4*C*C/ts[len(ts)//2]/1e9:.0f} TF")
