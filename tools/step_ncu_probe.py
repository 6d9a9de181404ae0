"""Three XL forwards (batch 1024) through the one-call product path, for ncu: the third is the one to capture
(`-s <2 x launches per forward> -c <launches per forward>`; the script prints the count).  python tools/step_ncu_probe.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import sdpnet_b200 as sdp  # noqa: E402
from bench import CONFIGS, NUM_REGISTERS  # noqa: E402

cfg, B0 = CONFIGS["XL"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else B0
torch.manual_seed(0)
model = sdp.MainModel.from_dict(**cfg).eval().to("cuda")
eng = model.engine()
x = torch.randn(B, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).bfloat16()
for i in range(3):
    sdp.ops.launch_count(reset=True)
    eng.forward(x, NUM_REGISTERS)
    torch.cuda.synchronize()
    print("forward", i, "launches", sdp.ops.launch_count(), flush=True)
