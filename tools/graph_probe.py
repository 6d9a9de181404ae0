"""CUDA-graph replay of the whole forward vs. stream launches (GPU box): does removing per-launch host work and
shortening kernel-to-kernel gaps pay at XL?  python tools/graph_probe.py [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpnet_b200 as sdp  # noqa: E402
from bench import CONFIGS, NUM_REGISTERS  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cfg, _ = CONFIGS["XL"]
    torch.manual_seed(0)
    eng = sdp.MainModel.from_dict(**cfg).eval().to("cuda").engine()
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1234)).cuda().bfloat16()
    for _ in range(3):
        ref = eng.forward(x, NUM_REGISTERS)
    torch.cuda.synchronize()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    t_stream = timed(lambda: eng.forward(x, NUM_REGISTERS))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.forward(x, NUM_REGISTERS)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = eng.forward(x, NUM_REGISTERS)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    ref2 = eng.forward(x, NUM_REGISTERS)
    torch.cuda.synchronize()
    print("ref is out:", ref.data_ptr() == out.data_ptr(), "ref2 is ref:", ref2.data_ptr() == ref.data_ptr(),
          "eager run-to-run max diff", float((ref2.float() - ref.float()).abs().max()),
          "graph vs eager max diff", float((out.float() - ref2.float()).abs().max()), "nan", int(torch.isnan(out.float()).sum()))
    g.replay()
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref2))
    t_graph = timed(g.replay)
    t_stream2 = timed(lambda: eng.forward(x, NUM_REGISTERS))
    eng.forward_graph(x, NUM_REGISTERS)                              # captures
    t_api = timed(lambda: eng.forward_graph(x, NUM_REGISTERS))       # the engine's own input copy + replay + output clone
    print(f"Engine.forward_graph {t_api:.2f} ms ({B / t_api * 1e3:.0f} img/s), identical: "
          f"{bool(torch.equal(eng.forward_graph(x, NUM_REGISTERS), ref2))}")
    print(f"B={B}: stream {t_stream:.2f} ms ({B / t_stream * 1e3:.0f} img/s) | graph replay {t_graph:.2f} ms "
          f"({B / t_graph * 1e3:.0f} img/s) | stream again {t_stream2:.2f} ms | graph output identical: {same}")


if __name__ == "__main__":
    main()
