"""N > 1 path on CPU: batch sharding + logits all_gather + count all_reduce over gloo, world_size 2.
The per-rank forward is replaced by the oracle on the rank's shard (CPU); the collective plumbing is
exactly what bench.py / an evaluation loop uses on NCCL."""
import importlib.util
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_shard():
    spec = importlib.util.spec_from_file_location("sdp_shard", os.path.join(ROOT, "sdp-net_b200", "shard.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sdpnet_oracle as O
    shard = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4, output_classes=7, max_image_size=[4, 4],
               head_output_from_register=True, conv_first=False)
    sd = O.synth_state_dict(cfg, seed=0, stress=True)
    x = torch.randn(n_total, 3, 16, 16, generator=torch.Generator().manual_seed(1234))
    labels = torch.arange(n_total) % 7
    lo, hi = shard.shard_range(n_total, rank, world)
    local = O.forward(sd, cfg, x[lo:hi], 3)
    full = shard.gather_logits(local, n_total)
    # EvalMeter's cross-rank reduction is a SUM all_reduce of (a clone of) its accumulators (CPU tensor here)
    acc4 = torch.tensor([float(rank + 1), 2.0, 3.0, float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(acc4, op=dist.ReduceOp.SUM)
    assert acc4.tolist() == [3.0, 4.0, 6.0, float(n_total)]
    correct, total = shard.reduce_counts(float((local.argmax(-1) == labels[lo:hi]).sum()), hi - lo)
    torch.save({"full": full, "correct": correct, "total": total, "lo": lo, "hi": hi}, f"{out_dir}/r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    shard = _load_shard()
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gather_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sdpnet_oracle as O
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_total = 5                                   # uneven: shards of 3 and 2
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    cfg = dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4, output_classes=7, max_image_size=[4, 4],
               head_output_from_register=True, conv_first=False)
    sd = O.synth_state_dict(cfg, seed=0, stress=True)
    x = torch.randn(n_total, 3, 16, 16, generator=torch.Generator().manual_seed(1234))
    ref = O.forward(sd, cfg, x, 3)
    labels = torch.arange(n_total) % 7
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 3, 3, 5)
    for r in (r0, r1):
        assert r["full"].shape == ref.shape
        assert (r["full"] - ref).abs().max() < 1e-5          # every rank sees all logits, in order
        assert r["total"] == n_total
        assert r["correct"] == float((ref.argmax(-1) == labels).sum())
