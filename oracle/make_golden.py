"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

For every case: weights come from `sdpnet_oracle.synth_state_dict(cfg, seed, stress)` and are
loaded into the reference `MainModel` with `load_state_dict(strict=True)` (which also proves
that the synthetic key layout is the reference's), the input is a seeded randn, and the
fixture stores the reference's logits / x_raw / registers / per-stage activations (forward
hooks) in fp32, plus a checksum of the weights so RNG drift is detected.  One tiny case
also embeds its weights so the oracle stays pinned even if torch's RNG stream changes.

The fixtures pin `oracle/sdpnet_oracle.py` (tests/test_oracle_golden.py) and are the
golden vectors the CUDA path is checked against (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import sdpnet_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TINY = dict(embedding_dim=32, n_head=2, num_blocks=2, patch_size=4, output_classes=10,
            max_image_size=[8, 8])
YAML_FLAGS = dict(conv_first=False, head_output_from_register=True, simple_mlp_output=False,
                  output_head_bias=False, normalize_qv=True, mixer_deptwise_bias=False,
                  mixer_ffn_bias=False, conv_embedding=False, activation="gelu",
                  embedding_activation="none", ff_multiplication_factor=4, conv_block_num=2,
                  max_num_registers=5)

CASES = {
    # name: (cfg, H, W, B, num_registers, seed, stress, embed_weights)
    "yaml_r4_refinit": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 3}, 16, 16, 3, 3, 0, False, True),
    "yaml_r5_stress": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 7}, 32, 32, 3, 4, 1, True, False),
    "yaml_r1_stress": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 5, "n_head": 4}, 32, 28, 2, 0, 2, True, False),
    "cifar_path": (dict(patch_size=2, embedding_dim=32, num_blocks=2, n_head=2,
                        conv_kernel_size=5, conv_embedding=True, max_image_size=[8, 8],
                        head_output_from_register=False, output_classes=10,
                        conv_embedding_kernel_size=5, embedding_activation="gelu"),
                   16, 16, 3, 3, 3, True, False),
    "biases_relu_nonsquare": (dict(embedding_dim=32, n_head=2, num_blocks=1, patch_size=4,
                                   output_classes=12, max_image_size=[6, 4], conv_kernel_size=3,
                                   activation="relu", embedding_activation="gelu",
                                   conv_first=True, head_output_from_register=True,
                                   simple_mlp_output=True, output_head_bias=True,
                                   normalize_qv=False, mixer_deptwise_bias=True,
                                   mixer_ffn_bias=True, conv_block_num=1,
                                   ff_multiplication_factor=2, max_num_registers=3),
                              16, 24, 2, 1, 4, True, False),
    "headbias_mlp": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 3, "output_head_bias": True,
                      "num_blocks": 1}, 16, 16, 2, 4, 5, True, False),
    "act_tanh": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 3, "activation": "tanh",
                  "num_blocks": 1}, 16, 16, 2, 3, 6, True, False),
    "act_sigmoid": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 3, "activation": "sigmoid",
                     "num_blocks": 1}, 16, 16, 2, 3, 7, True, False),
    "act_leaky_selu": ({**TINY, **YAML_FLAGS, "conv_kernel_size": 3, "activation": "leaky_relu",
                        "embedding_activation": "selu", "num_blocks": 1}, 16, 16, 2, 3, 8, True, False),
    # BASELINE.json configs[0], README-row reading: XXS on a 32x32 CIFAR-shaped batch
    "xxs_readme_32": (dict(embedding_dim=128, num_blocks=7, n_head=4, patch_size=16,
                           conv_kernel_size=7, output_classes=100, max_image_size=[16, 16],
                           **{k: v for k, v in YAML_FLAGS.items()}),
                      32, 32, 8, 4, 9, False, False),
    # BASELINE.json configs[0], literal cifar100_test.py:65-91 constructor
    "xxs_cifar_ctor": (dict(conv_first=True, max_image_size=[16, 16], patch_size=2,
                            embedding_dim=128, num_blocks=16, n_head=4, activation="gelu",
                            embedding_activation="none", conv_kernel_size=5, conv_block_num=2,
                            ffn_dropout=0.1, attn_dropout=0.1, output_classes=100,
                            max_num_registers=5, ff_multiplication_factor=4,
                            head_output_from_register=False, simple_mlp_output=False,
                            output_head_bias=True, normalize_qv=True, mixer_deptwise_bias=True,
                            mixer_ffn_bias=True, stochastic_depth_p=[0.05, 0.05],
                            conv_embedding=True, conv_embedding_kernel_size=5),
                       32, 32, 2, 3, 10, False, False),
}


def checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def run_reference(cfg, sd, x, num_registers):
    from model import MainModel
    from layers import ConvMixer, EncoderLayer

    torch.set_float32_matmul_precision("highest")
    model = MainModel.from_dict(**cfg).eval()
    model.load_state_dict(sd, strict=True)
    stages, order = {}, []
    state = {"reg": None}

    def tok(xs, rs):
        return torch.cat([rs, xs.flatten(2).transpose(1, 2)], 1).detach().clone()

    def emb_hook(_m, _i, out):
        state["reg"] = out[1]
        stages["embed"] = tok(out[0], out[1])

    def enc_hook(name):
        def h(_m, _i, out):
            state["reg"] = out[1]
            stages[name] = tok(out[0], out[1])
        return h

    def mix_hook(name):
        def h(_m, _i, out):
            stages[name] = tok(out, state["reg"])
        return h

    model.embedding_layer.register_forward_hook(emb_hook)
    for i, blk in enumerate(model.blocks):
        blk.t_block.register_forward_hook(enc_hook(f"b{i}.enc"))
        for j, mx in enumerate(blk.conv_blocks):
            mx.register_forward_hook(mix_hook(f"b{i}.mixer{j}"))
    model.final_block.t_block.register_forward_hook(enc_hook("final"))
    with torch.no_grad():
        logits, x_raw, reg = model(x.clone(), num_registers, True)
    return logits, x_raw, reg, stages


def kelu_layer_case():
    """Layer-level KeLU fixture (unreachable through MainModel):
    EncoderLayer(activation_func=KeLu), training_utilities.py:91-92 + layers.py:308."""
    from layers import EncoderLayer
    from training_utilities import KeLu

    torch.set_float32_matmul_precision("highest")
    g = torch.Generator().manual_seed(77)
    C, nh = 32, 4
    layer = EncoderLayer(embedding_dim=C, n_head=nh, activation_func=KeLu,
                         multiplication_factor=4).eval()
    sd = {}
    for k, v in layer.state_dict().items():
        if k.endswith("weight") and v.ndim == 2:
            sd[k] = torch.randn(v.shape, generator=g) / np.sqrt(v.shape[1])
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1 + 0.25 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.3 * torch.randn(v.shape, generator=g)
    # large ff bias spread so |pre-activation| crosses +-3.5 (all three KeLU branches)
    sd["ff_linear1.bias"] = 3.0 * torch.randn(sd["ff_linear1.bias"].shape, generator=g)
    layer.load_state_dict(sd, strict=True)
    x = torch.randn(2, C, 4, 4, generator=g)
    reg = torch.randn(2, 3, C, generator=g)
    with torch.no_grad():
        xo, ro = layer(x, reg)
    out = {f"sd/{k}": v.numpy() for k, v in sd.items()}
    out.update(x=x.numpy(), reg=reg.numpy(), x_out=xo.numpy(), reg_out=ro.numpy(),
               meta=np.array(json.dumps(dict(embedding_dim=C, n_head=nh, activation="kelu"))))
    np.savez_compressed(os.path.join(OUT, "layer_encoder_kelu.npz"), **out)
    # and the bare function on a grid, for the elementwise epilogue
    t = torch.linspace(-6, 6, 4001)
    np.savez_compressed(os.path.join(OUT, "act_kelu_grid.npz"), x=t.numpy(), y=KeLu(t).numpy())


def metrics_case():
    """Reference losses / accuracy (model_test.py:70-83) on seeded logits."""
    from training_utilities import BCEWithLogitsLoss
    g = torch.Generator().manual_seed(99)
    logits = 3.0 * torch.randn(37, 100, generator=g)
    labels = torch.randint(0, 100, (37,), generator=g)
    logits[torch.arange(0, 37, 3), labels[::3]] += 8.0        # make a third of the rows correct
    out = dict(logits=logits.numpy(), labels=labels.numpy())
    out["ce"] = np.array(float(torch.nn.CrossEntropyLoss()(logits, labels)))
    out["acc"] = np.array(float((logits.argmax(1) == labels).float().mean()))
    for ls in (0.0, 0.1):
        out[f"bce_ls{ls}"] = np.array(float(BCEWithLogitsLoss(num_classes=100, label_smoothing=ls)(logits, labels)))
    np.savez_compressed(os.path.join(OUT, "act_eval_metrics.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (cfg, H, W, B, nr, seed, stress, embed) in CASES.items():
        sd = O.synth_state_dict(cfg, seed=seed, stress=stress)
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(1234))
        logits, x_raw, reg, stages = run_reference(cfg, sd, x, nr)
        # self-check of the restatement at generation time
        st2 = {}
        lo, xo, ro = O.forward(sd, cfg, x, nr, True, stages=st2)
        err = max(float((lo - logits).abs().max()), float((xo - x_raw).abs().max()),
                  float((ro - reg).abs().max()))
        meta = dict(cfg=cfg, H=H, W=W, B=B, num_registers=nr, seed=seed, stress=stress,
                    checksum=checksum(sd), torch=torch.__version__, input_seed=1234)
        out = dict(meta=np.array(json.dumps(meta)), logits=logits.numpy(), x_raw=x_raw.numpy(),
                   registers=reg.contiguous().numpy())
        total = sum(v.numel() for v in stages.values())
        keep = stages if total < 400_000 \
            else {k: v for k, v in stages.items() if k in ("embed", "b0.enc", "b0.mixer0", "b0.mixer1", "final")}
        for k, v in keep.items():
            out["stage/" + k] = v.numpy()
        if embed:
            for k, v in sd.items():
                out["sd/" + k] = v.numpy()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:24s} oracle-vs-reference max|d|={err:.2e}  stages={len(keep)} "
              f"size={os.path.getsize(path) / 1024:.0f} KiB")
    kelu_layer_case()
    metrics_case()
    print("done ->", OUT)


if __name__ == "__main__":
    main()
