#!/usr/bin/env python
"""Benchmark of the SdP-Net forward path (BASELINE.json metric: XL 224^2 bf16 forward images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config XL|M|S] [--batch B]
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference forward)

One JSON line on stdout (rank 0).  A "step" is one forward over one batch of B synthetic images per
GPU.  `value` = images/s with the batch already resident in HBM (device-timed with CUDA events,
max over ranks); `e2e` = the same through the public module API with HOST inputs (pinned fp32
images copied H2D and logits read back D2H inside the timed region); `roofline` = the dominant
kernel family (the tcgen05 GEMM), timed live with CUDA events in an instrumented pass of the same
forward; `cpu_baseline` = the oracle (a port of the reference's PyTorch forward) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

YAML = dict(n_head=8, activation="gelu", embedding_activation="none", conv_kernel_size=7, output_classes=1000,
            conv_block_num=2, ff_multiplication_factor=4, max_image_size=[16, 16], max_num_registers=5,
            conv_first=False, head_output_from_register=True, simple_mlp_output=False, output_head_bias=False,
            normalize_qv=True, mixer_deptwise_bias=False, mixer_ffn_bias=False, conv_embedding=False)
CONFIGS = {   # BASELINE.json configs[1..3]
    "S": (dict(YAML, embedding_dim=512, num_blocks=12, patch_size=16), 256),
    "M": (dict(YAML, embedding_dim=768, num_blocks=12, patch_size=16), 512),
    "XL": (dict(YAML, embedding_dim=768, num_blocks=17, patch_size=14), 1024),
}
NUM_REGISTERS = 4   # R = 5
METRIC = "SdP-Net XL 224^2 bf16 fwd images/sec"


def workload_name(size, batch):
    cfg = CONFIGS[size][0]
    return (f"SdP-Net {size} ({cfg['num_blocks']} blocks, embed {cfg['embedding_dim']}, patch {cfg['patch_size']}, "
            f"conv 7, 8 heads, R=5) 224^2 eval forward, batch {batch} per GPU, random-init weights (seed 0)")


def gemm_traffic(config, batch):
    """DRAM bytes per GEMM launch (dram__bytes_read + dram__bytes_write, averaged over the step's GEMM launches) from
    the committed `ncu --set full` capture; None when the capture was taken on another workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")) as f:
            t = json.load(f)
        if t.get("config") != config or int(t.get("batch", 0)) != int(batch):
            return None
        return t["avg_dram_gb_per_launch"] * 1e9
    except (OSError, ValueError, KeyError):
        return None


def gemm_bytes_per_step(cfg, M):
    """Algorithmic bytes of the step's big GEMMs: A + W + out in bf16; the GEMMs that write the residual stream read and
    write both of its planes (hi + lo, 2 bytes each)."""
    C, F = int(cfg["embedding_dim"]), int(cfg.get("ff_multiplication_factor", 4)) * int(cfg["embedding_dim"])
    def g(K, N, res):
        return 2 * (M * K + N * K + M * N * (4 if res else 1))
    nb, cb = int(cfg["num_blocks"]), int(cfg["conv_block_num"])
    enc = g(C, 3 * C, False) + g(C, C, True) + g(C, F, False) + g(F, C, True)
    mix = g(C, C, True) + g(C, 4 * C, False) + g(4 * C, C, True)
    return (nb + 1) * enc + nb * cb * mix


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], src="measured")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [s.strip() for s in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        # keep the samples taken under load (upper half) for the median
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_forward_fn(cfg):
    """(callable x -> logits, kind, description): the reference's own `MainModel` (fp32, eval, CPU) from the vendored
    baseline/_ref copy -- kind "reference" -- or, when no copy travelled to this box, the oracle port of it."""
    import torch
    import sdpnet_oracle as O
    import reference_loader as RL
    torch.set_float32_matmul_precision("highest")
    sd = O.synth_state_dict(cfg, seed=0)
    if RL.available():
        model = RL.build_model(cfg, sd)
        torch.set_float32_matmul_precision("highest")
        return (lambda x: model(x, NUM_REGISTERS)), "reference", "unmodified reference MainModel (baseline/_ref/model.py), torch CPU"
    return (lambda x: O.forward(sd, cfg, x, NUM_REGISTERS)), "port", "oracle/sdpnet_oracle.py (port of the reference forward), torch CPU"


def cpu_forward_rate(cfg, batch, warm_batch=2, threads=None, min_seconds=12.0, max_seconds=30.0):
    """images/s of the reference forward on the host cores, fp32: forwards of `batch` images repeated until
    `min_seconds` of CPU work have been timed (a bounded sample of the workload)."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    fwd, kind, what = cpu_forward_fn(cfg)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        fwd(torch.randn(warm_batch, 3, 224, 224, generator=g))
        x = torch.randn(batch, 3, 224, 224, generator=g)
        n, t0 = 0, time.perf_counter()
        while True:
            fwd(x)
            n += batch
            dt = time.perf_counter() - t0
            if dt >= min_seconds or dt + dt / (n // batch) > max_seconds:
                break
    return n / dt, threads, dt, n, kind, what


def run_reference(args):
    """--impl reference: the reference's own CPU forward -- the UNMODIFIED `MainModel` from baseline/_ref (vendored by
    tools/vendor_reference.sh; falls back to the oracle port if that copy is absent) -- on all host cores, a bounded
    sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cfg, _ = CONFIGS[args.config]
    sample = args.cpu_batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fwd, kind, what = cpu_forward_fn(cfg)
    x = torch.randn(sample, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(x[:2])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(x)
        dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    batch = args.batch or CONFIGS[args.config][1]
    line = {
        "impl": "reference", "metric": METRIC if args.config == "XL" else f"SdP-Net {args.config} 224^2 fwd images/sec",
        "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, batch), "num_registers": NUM_REGISTERS},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} images per step (fp32, {what})"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_comparators(cfg, state_dict, x_host, dev, engine_logits, steps=3, warmup=2, compile_budget_s=420.0):
    """The reference ITSELF on this GPU: the unmodified `MainModel` from baseline/_ref
    with the engine's weights and batch, (a) eager under `torch.autocast('cuda', bfloat16)` -- its training /
    inference precision path, training_tools.py:85-86 -- and (b) `torch.compile(model)` in eval, model_test.py:16,64
    (fp32 parameters, TF32 matmuls as model.py:9 sets them, no autocast: exactly that script), plus (c) compile under
    autocast, the fastest stock configuration.  PyTorch library kernels only; none of this repo's code runs here."""
    import torch
    import reference_loader as RL
    out = {"unit": "images/s", "batch": int(x_host.shape[0]), "steps": steps, "warmup": warmup}
    if not RL.available():
        out["unavailable"] = "baseline/_ref not vendored (tools/vendor_reference.sh)"
        return out
    prec = torch.get_float32_matmul_precision()
    model = RL.build_model(cfg, {k: v.detach().cpu() for k, v in state_dict.items()}).to(dev)
    torch.set_float32_matmul_precision("high")          # what `import model` leaves behind (model.py:9)
    x = x_host.to(dev)                                    # fp32 images, as the reference's loaders emit
    B = x.shape[0]

    def timed(fn):
        with torch.no_grad():
            for _ in range(warmup):
                y = fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                y = fn()
            e1.record()
            torch.cuda.synchronize(dev)
        return B * steps / (e0.elapsed_time(e1) / 1e3), y

    def eager_autocast():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x, NUM_REGISTERS)

    try:
        v, y = timed(eager_autocast)
        out["eager_autocast_bf16"] = v
        y = y.float()
        out["engine_vs_reference_bf16_gpu"] = {
            "logits_max_abs": float((engine_logits.float() - y).abs().max()),
            "top1_agreement": float((engine_logits.argmax(-1) == y.argmax(-1)).float().mean())}
        v, _ = timed(lambda: model(x, NUM_REGISTERS))
        out["eager_fp32_tf32"] = v
    except Exception as e:      # noqa: BLE001 -- a comparator must never take the engine line down
        out["eager_error"] = repr(e)[:200]
    try:
        t0 = time.perf_counter()
        cmodel = torch.compile(model)
        v, _ = timed(lambda: cmodel(x, NUM_REGISTERS))
        out["torch_compile"] = v
        out["torch_compile_seconds_incl_compile"] = time.perf_counter() - t0
        if time.perf_counter() - t0 < compile_budget_s:
            def compiled_autocast():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return cmodel(x, NUM_REGISTERS)
            v, _ = timed(compiled_autocast)
            out["torch_compile_autocast_bf16"] = v
    except Exception as e:      # noqa: BLE001
        out["torch_compile_error"] = repr(e)[:300]
    torch.set_float32_matmul_precision(prec)
    del model
    torch.cuda.empty_cache()
    return out


def cublas_yardstick(cfg, M, dev, seconds=1.0):
    """What the vendor library reaches on THIS step's GEMM shapes at this M, in the same power-capped regime (a yardstick
    only: cuBLAS is not on the product path).  The 'sustained peak' in MEASURED_PEAKS.json is an 8192^3 number whose
    operands live in L2; at M = B*S with K = 768 / 3072 the library itself stays well below it."""
    import torch
    C, F = int(cfg["embedding_dim"]), int(cfg["ff_multiplication_factor"]) * int(cfg["embedding_dim"])
    nb, cb = int(cfg["num_blocks"]), int(cfg["conv_block_num"])
    shapes = [(3 * C, C, nb + 1), (C, C, nb + 1 + nb * cb), (F, C, nb + 1), (C, F, nb + 1), (4 * C, C, nb * cb), (C, 4 * C, nb * cb)]
    tot_ms, tot_fl, per = 0.0, 0.0, {}
    for N, K, count in shapes:
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.matmul(A, W.t(), out=out)
        e1.record()
        torch.cuda.synchronize(dev)
        reps = max(12, int(seconds * 1e3 / max(e0.elapsed_time(e1) / 3, 1e-3)))     # ~`seconds` of back-to-back launches:
        e0.record()                                                                  # the power-capped regime of the step
        for _ in range(reps):
            torch.matmul(A, W.t(), out=out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        per[f"N{N} K{K}"] = round(2.0 * M * N * K / ms / 1e9, 1)
        tot_ms += ms * count
        tot_fl += 2.0 * M * N * K * count
        del A, W, out
    return {"tflops_launch_weighted": tot_fl / tot_ms / 1e9, "ms_per_step_if_every_gemm_were_plain_cublas": tot_ms, "by_shape_tflops": per,
            "how": f"torch.matmul (cuBLAS) bf16, ~{seconds:.1f} s of back-to-back launches per shape (power-capped like the step), "
                   "no epilogue work at all"}


def algorithmic_flops_per_image(cfg, T, R):
    """2 * MAC of one forward, SURVEY.md §8(d): patcher + nb * (cbn * mixer + encoder) + final encoder + head; the
    mixers are counted on the T patch tokens, LN / GELU / softmax flops are not counted."""
    C, m, nb, cbn, K, p, k = (cfg["embedding_dim"], cfg["ff_multiplication_factor"], cfg["num_blocks"], cfg["conv_block_num"],
                              cfg["output_classes"], cfg["patch_size"], cfg["conv_kernel_size"])
    S = T + R
    patch = 2 * T * C * 3 * p * p
    mixer = 2 * T * C * k * k + 2 * T * C * C + 16 * T * C * C
    enc = 6 * S * C * C + 4 * S * S * C + 2 * S * C * C + 4 * m * S * C * C
    head = 2 * C * K + 2 * K * K if cfg["head_output_from_register"] and not cfg["simple_mlp_output"] else 2 * C * K
    return float(patch + nb * (cbn * mixer + enc) + enc + head)


def gemm_flops_per_image(cfg, T, R):
    C, m, nb, cbn, K, p = (cfg["embedding_dim"], cfg["ff_multiplication_factor"], cfg["num_blocks"],
                           cfg["conv_block_num"], cfg["output_classes"], cfg["patch_size"])
    S = T + R
    enc = 6 * S * C * C + 2 * S * C * C + 4 * m * S * C * C
    mixer = 2 * S * C * C + 16 * S * C * C          # GEMMs run over all S rows (register rows masked)
    patch = 2 * T * C * 3 * p * p
    head = 2 * C * K + 2 * K * K
    return patch + nb * (cbn * mixer + enc) + enc + head


def profile_families(eng, x, R, steps):
    """Per-kernel-family device time of one forward (CUDA events around every op of the op-by-op
    sequence -- same kernels, same order as sdp_forward)."""
    import torch
    import sdpnet_b200 as sdp
    ops = sdp.ops
    fam = {}
    pending = []
    orig = {}

    def gemm_kind(a, k):
        """GEMM launches by shape and epilogue: N x K + what rides in the epilogue."""
        W = a[1]
        tag = f"gemm N{W.shape[0]} K{k.get('K') or W.shape[1]}"
        if k.get("ln_fold") is not None:
            tag += " +lnfold"
        if k.get("headnorm") is not None:
            tag += " +headnorm"
        if k.get("act", "none") not in ("none", 0):
            tag += f" +{k['act']}"
        if k.get("residual") is not None:
            tag += " +residual" + ("(hi+lo)" if k.get("out_lo") is not None else "")
        if k.get("stats_out") is not None:
            tag += " +stats"
        return tag

    def wrap(name, family):
        fn = getattr(ops, name)
        orig[name] = fn

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            pending.append((family, gemm_kind(a, k) if name == "gemm" else name, e0, e1))
            return r
        setattr(ops, name, timed)

    for name, family in [("gemm", "gemm_bf16_tc"), ("layernorm_rows", "layernorm_rows"), ("ln_dwconv", "ln_dwconv"), ("ln_dwconv_slab", "ln_dwconv"),
                         ("attention", "attention"), ("im2col_patches", "other"), ("fill_registers", "other"),
                         ("pool_ln", "other"), ("row_stats", "other")]:
        wrap(name, family)
    try:
        for _ in range(steps):
            eng.forward(x, R - 1, staged=True)
        torch.cuda.synchronize()
    finally:
        for name, fn in orig.items():
            setattr(ops, name, fn)
    counts, detail = {}, {}
    for family, kind, e0, e1 in pending:
        ms = e0.elapsed_time(e1)
        fam[family] = fam.get(family, 0.0) + ms
        counts[family] = counts.get(family, 0) + 1
        d = detail.setdefault(kind, [0.0, 0])
        d[0] += ms
        d[1] += 1
    out = {k: {"ms_per_step": v / steps, "launches_per_step": counts[k] // steps} for k, v in fam.items()}
    out["_detail"] = {k: {"ms_per_launch": round(v[0] / v[1], 4), "launches_per_step": v[1] // steps,
                          "ms_per_step": round(v[0] / steps, 3)} for k, v in sorted(detail.items(), key=lambda kv: -kv[1][0])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="XL", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: BASELINE batch)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-comparators", action="store_true", help="skip the reference-on-this-GPU comparators")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import sdpnet_b200 as sdp          # the engine arm never touches oracle/ (only cpu_forward_rate / run_reference do)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left exactly as the caller set it (INFO to count ranks, or unset = silent): with it unset NCCL
        # prints nothing and stdout stays the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    cfg, default_batch = CONFIGS[args.config]
    B = args.batch or default_batch
    R = NUM_REGISTERS + 1
    p = cfg["patch_size"]
    T = (224 // p) ** 2

    # weights: random init of that architecture through the module API (the reference's own initialisation,
    # model.py:121-126, seed 0 on every rank) -> packed engine
    torch.manual_seed(0)
    model = sdp.MainModel.from_dict(**cfg).eval().to(dev)
    eng = model.engine()

    # this rank's shard of the synthetic batch: host (pinned, fp32 like the reference's loaders) + device bf16
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True).bfloat16()
    x_stage = torch.empty_like(x_host, device=dev)
    logits_host = torch.empty(B, cfg["output_classes"], dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ranks_seen = 1
    if world > 1:                                        # one SUM all_reduce of ones: every rank of the job took part
        ones = torch.ones(1, device=dev)
        dist.all_reduce(ones)
        ranks_seen = int(ones.item())

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------------
    for _ in range(args.warmup):
        eng.forward(x_dev, NUM_REGISTERS)
    barrier()
    sdp.ops.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            logits = eng.forward(x_dev, NUM_REGISTERS)
        e1.record()
        barrier()
    launches = sdp.ops.launch_count()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end through the module API, host buffers in and out -------------------------
    # Every step copies ITS OWN images from pinned host memory and reads ITS logits back, inside the timed
    # region; the copy of step i+1 runs on a second stream under the forward of step i (two staging buffers).
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    stages = [x_stage, torch.empty_like(x_stage)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def h2d(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the forward that last read this buffer is done
            stages[i % 2].copy_(x_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(main_stream)
        h2d(0)
        for i in range(n):
            if i + 1 < n:
                h2d(i + 1)
            main_stream.wait_event(ready[i % 2])
            out = model(stages[i % 2], NUM_REGISTERS)
            consumed[i % 2].record(main_stream)
            logits_host.copy_(out, non_blocking=True)

    e2e_run(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_run(args.steps)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    e2e = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "images/s",
           "h2d_bytes_per_step": x_host.numel() * 4 * world, "d2h_bytes_per_step": logits_host.numel() * 4 * world,
           "api": "sdpnet_b200.MainModel.__call__ (one sdp_forward C-ABI call); pinned fp32 host images, H2D of "
                  "step i+1 overlapped with the forward of step i on a copy stream; logits D2H every step"}

    # ---- off the timed path: gather logits + 2-scalar reduction (training_utilities.py:33,72-73) ----
    if world > 1:
        gathered = [torch.empty_like(logits) for _ in range(world)]
        dist.all_gather(gathered, logits)
        stats = torch.tensor([float(logits.argmax(-1).eq(0).sum()), float(B)], device=dev)
        dist.all_reduce(stats)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops_img = algorithmic_flops_per_image(cfg, T, R)
    line = {
        "metric": METRIC if args.config == "XL" else f"SdP-Net {args.config} 224^2 bf16 fwd images/sec",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args.config, B), "global_batch": B * world, "num_registers": NUM_REGISTERS,
                   "parallelism": f"dp{world} (batch-sharded, no collective on the timed path)",
                   "l2": "no explicit flush: per-step inputs (%d MB) and activations (%.1f GB) far exceed the 126 MB L2"
                         % (x_dev.numel() * 2 >> 20, eng.buffers(B, 224 // p, 224 // p, R).nbytes() / 2 ** 30)},
        "clocks": clk.summary(),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "comm": {"backend": "nccl" if world > 1 else None, "world_size": world, "ranks_in_all_reduce": ranks_seen,
                 "collectives_on_timed_path": 0},
        "model_tflops": value / world * flops_img / 1e12,
        "model_frac_of_sustained_peak": value / world * flops_img / 1e12 / peaks["sustained"],
        "algorithmic_gflop_per_image": flops_img / 1e9,
    }

    if not args.no_profile:
        fam = profile_families(eng, x_dev, R, 2)
        detail = fam.pop("_detail")
        gf = gemm_flops_per_image(cfg, T, R) * B
        gms = fam["gemm_bf16_tc"]["ms_per_step"]
        nl = fam["gemm_bf16_tc"]["launches_per_step"]
        achieved = gf / (gms / 1e3) / 1e12
        line["roofline"] = {
            "kernel": "gemm_bf16_tc_kernel (tcgen05/TMEM/TMA)", "bound": "tensor", "achieved": achieved,
            "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["sustained"],
            "traffic": gemm_traffic(args.config, B),
            "traffic_how": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 177 GEMM launches of one "
                           "whole forward captured with ncu --set full (profiles/r02_gemm_traffic.json, keyed by kernel "
                           "name; tools/step_ncu_probe.py + tools/step_ncu_summary.py); compare algorithmic_bytes_per_launch",
            "algorithmic_bytes_per_launch": gemm_bytes_per_step(cfg, B * (T + R)) / nl,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
            "launches_per_step": nl, "avg_launch_ms": gms / nl, "avg_launch_gflop": gf / nl / 1e9,
            "share_of_step": gms / sum(v["ms_per_step"] for v in fam.values()),
            "how": "CUDA events around every launch of an instrumented op-by-op pass of the same forward",
        }
        if world == 1 and not args.no_comparators:
            y = cublas_yardstick(cfg, B * (T + R), dev)
            line["roofline"]["cublas_same_shapes"] = y
            line["roofline"]["frac_of_cublas_same_shapes"] = achieved / y["tflops_launch_weighted"]
            line["roofline"]["note"] = ("the GEMM family carries, in its epilogues, every token LayerNorm of the model (folded), the GELUs, the "
                                        "per-head q/k LayerNorm and the two-plane (hi + lo) residual stream; cublas_same_shapes is the bare "
                                        "library GEMM on the same shapes with none of that")
        line["kernel_families_ms_per_step"] = {k: round(v["ms_per_step"], 3) for k, v in fam.items()}
        line["kernel_detail"] = detail
        # bandwidth-bound families, for the record (algorithmic bytes: read + write of [B,S,C] bf16)
        S = T + R
        if "ln_dwconv" in fam:
            nbytes = 2 * B * T * cfg["embedding_dim"] * 2 * cfg["num_blocks"] * cfg["conv_block_num"]
            line["ln_dwconv_gbs"] = nbytes / (fam["ln_dwconv"]["ms_per_step"] / 1e3) / 1e9
        if "layernorm_rows" in fam:
            nbytes = 2 * B * S * cfg["embedding_dim"] * 2 * fam["layernorm_rows"]["launches_per_step"]
            line["layernorm_gbs"] = nbytes / (fam["layernorm_rows"]["ms_per_step"] / 1e3) / 1e9
        line["hbm_peak_gbs"] = peaks["hbm"]

    if not args.no_comparators and world == 1:
        # the reference itself on this GPU, same weights and batch (eager autocast bf16, torch.compile)
        line["gpu_comparators"] = gpu_comparators(cfg, model.state_dict(), x_host, dev, logits)
        ec = line["gpu_comparators"].get("eager_autocast_bf16")
        if ec:
            best = max(v for k, v in line["gpu_comparators"].items()
                       if k in ("eager_autocast_bf16", "eager_fp32_tf32", "torch_compile", "torch_compile_autocast_bf16"))
            line["gpu_comparators"]["engine_over_eager_autocast"] = value / ec
            line["gpu_comparators"]["engine_over_best_reference"] = value / best

    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (host cores are shared by the ranks)
        v, cores, dt, n, kind, what = cpu_forward_rate(cfg, args.cpu_batch)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": kind,
                                "sample": f"{n} images in fp32 forwards of {args.cpu_batch} through the {what}, "
                                          f"{dt:.1f} s after a 2-image warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
