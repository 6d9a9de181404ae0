"""TEST INFRASTRUCTURE -- generates checkpoint fixtures with the REFERENCE'S OWN writers (run in the build container,
where /root/reference exists; the files it writes are committed under tests/golden/):

  ckpt_save_model.pt   SdPModel.save_model                 utility_layers.py:185-198   {"state_dict", "config"}
  ckpt_trainer.pt      Trainer._save_checkpoint            training_tools.py:203-226   {"model_state_dict" (with the
                       "module." prefix of the DDP wrapper, training_tools.py:36), "model_config", "optimizer_state",
                       "scheduler_state", "epoch"}
  ckpt_ema.pt          EMA_model.save_ema_model            training_tools.py:282-302   bare state_dict
  ckpt_expected.npz    the reference's fp32 CPU forward of the saved model (and of the EMA weights) on a seeded input

python oracle/make_golden_checkpoints.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
OUT = os.path.join(ROOT, "tests", "golden")

import reference_loader as RL  # noqa: E402
import sdpnet_oracle as O  # noqa: E402

CFG = dict(embedding_dim=32, num_blocks=2, n_head=2, activation="gelu", conv_kernel_size=5, patch_size=4,
           output_classes=11, conv_block_num=2, ff_multiplication_factor=4, max_image_size=[8, 8], max_num_registers=5,
           conv_first=False, head_output_from_register=True, simple_mlp_output=False, normalize_qv=True)


class _DDPLike(torch.nn.Module):
    """What DistributedDataParallel does to the key names: everything sits under `module.`."""

    def __init__(self, m):
        super().__init__()
        self.module = m


def main():
    torch.manual_seed(0)
    ref_model_mod = RL.load()
    sys.path.insert(0, "/root/reference")  # training_tools.py is not part of the vendored forward-only copy
    import training_tools as TT            # the reference's Trainer / EMA_model
    torch.set_float32_matmul_precision("highest")
    model = ref_model_mod.MainModel.from_dict(**CFG).eval()
    model.load_state_dict(O.synth_state_dict(CFG, seed=21, stress=True), strict=True)
    x = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        logits, x_raw, reg = model(x, 3, True)

    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        # (a) SdPModel.save_model writes "<name>.pt"
        model.save_model("ckpt_save_model")
        # (b) Trainer._save_checkpoint, run unbound on the handful of attributes it reads
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
        fake = types.SimpleNamespace(snapshot_dir=OUT, snapshot_name="ckpt_trainer.pt", model=_DDPLike(model),
                                     model_config=dict(CFG), optimizer=opt,
                                     scheduler=torch.optim.lr_scheduler.ConstantLR(opt, factor=1.0, total_iters=1), epoch=7)
        TT.Trainer._save_checkpoint(fake)
        # (c) EMA weights: a few updates towards a perturbed model so that they differ from the model's own
        ema = TT.EMA_model(model, decay=0.5, ema_model_name="ckpt_ema.pt")
        pert = ref_model_mod.MainModel.from_dict(**CFG).eval()
        pert.load_state_dict(O.synth_state_dict(CFG, seed=22, stress=True), strict=True)
        ema.update_parameters(pert)
        ema.save_ema_model()
    finally:
        os.chdir(cwd)
    ema_model = ref_model_mod.MainModel.from_dict(**CFG).eval()
    ema_model.load_state_dict(torch.load(os.path.join(OUT, "ckpt_ema.pt")), strict=True)
    with torch.no_grad():
        e_logits, e_raw, e_reg = ema_model(x, 3, True)
    np.savez_compressed(os.path.join(OUT, "ckpt_expected.npz"), x=x.numpy(), logits=logits.numpy(), x_raw=x_raw.numpy(),
                        registers=reg.numpy(), ema_logits=e_logits.numpy(), ema_x_raw=e_raw.numpy())
    for f in ("ckpt_save_model.pt", "ckpt_trainer.pt", "ckpt_ema.pt", "ckpt_expected.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
