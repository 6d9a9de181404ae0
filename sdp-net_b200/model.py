"""`MainModel` with the reference's constructor, `from_dict`, `state_dict` layout and
`forward(x, num_registers=3, return_raw_outputs=False)` (reference model.py:28-149,
utility_layers.py:94-198), running on the sdpnet_b200 CUDA engine.

The module tree exists to own the parameters under the reference's names; `forward` packs them
once (`Engine`) and runs the whole network through ONE `sdp_forward` C-ABI call in token-major
layout.  Eval only; CUDA only; bf16 by default, `set_precision('fp32')` for the 1e-4 mode.
"""
from __future__ import annotations

from typing import Callable

import torch
from numpy import arccos, cos
from torch import nn

from .engine import Engine, _act_name
from .layers import (Block, ClassificationHead, ConvEmbedding, ConvPatcher, EmbeddingLayer, FinalBlock,
                     _KernelModule)

# reference model.py:13-24.  Values are only ever inspected for their kind (engine._act_name).
activations = {
    "relu": nn.ReLU(), "gelu": nn.GELU(), "tanh": nn.Tanh(), "sigmoid": nn.Sigmoid(),
    "leaky_relu": nn.LeakyReLU(), "selu": nn.SELU(), "none": nn.Identity(), "gelu_tanh": nn.GELU("tanh"),
}


class SdPModel(_KernelModule):
    """reference utility_layers.py:94-198 (config carrying, from_dict / from_pretrained / save_model)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.config = {}

    def return_num_params(self) -> dict:
        tr = sum(p.numel() for p in self.parameters() if p.requires_grad)
        nt = sum(p.numel() for p in self.parameters() if not p.requires_grad)
        return {"Trainable_params": tr, "Non_trainable_params": nt}

    @classmethod
    def from_dict(cls, **kwargs):
        model = cls(**kwargs)
        model.config = kwargs
        return model

    @classmethod
    def from_pretrained(cls, file_name, map_location="cpu"):
        blob = torch.load(file_name, map_location=map_location)
        model = cls.from_dict(**blob["config"])
        model.load_state_dict(blob["state_dict"])
        return model

    def save_model(self, file_name=None):
        fn = "Model" if file_name is None else file_name
        torch.save({"state_dict": self.state_dict(), "config": self.config}, f"{fn}.pt")


class MainModel(SdPModel):
    def __init__(self, embedding_dim: int = 128, num_blocks: int = 10, n_head: int = 4,
                 activation: Callable = "gelu", conv_kernel_size: int = 5, patch_size: int = 16,
                 ffn_dropout: float = 0.2, attn_dropout: float = 0.2, output_classes: int = 1000,
                 conv_block_num: int = 2, ff_multiplication_factor: int = 4, max_image_size=[14, 14],
                 max_num_registers: int = 5, embedding_activation: Callable = "none", conv_first: bool = True,
                 head_output_from_register: bool = False, simple_mlp_output: bool = False,
                 output_head_bias: bool = False, normalize_qv: bool = True, stochastic_depth_p=[0.0, 0.0],
                 mixer_deptwise_bias: bool = False, mixer_ffn_bias: bool = False, fast_att: bool = True,
                 conv_embedding: bool = False, conv_embedding_kernel_size: int = 5):
        super().__init__()
        if isinstance(activation, str) and activation.lower() == "kelu":
            # upstream: KeLu is a bare function and nn.Sequential rejects it
            raise TypeError("KeLu is not a Module subclass")
        self._engine_cfg = dict(
            embedding_dim=embedding_dim, num_blocks=num_blocks, n_head=n_head, activation=_act_name(activation),
            conv_kernel_size=conv_kernel_size, patch_size=patch_size, output_classes=output_classes,
            conv_block_num=conv_block_num, ff_multiplication_factor=ff_multiplication_factor,
            max_image_size=list(max_image_size), max_num_registers=max_num_registers,
            embedding_activation=_act_name(embedding_activation), conv_first=conv_first,
            head_output_from_register=head_output_from_register, simple_mlp_output=simple_mlp_output,
            output_head_bias=output_head_bias, normalize_qv=normalize_qv, mixer_deptwise_bias=mixer_deptwise_bias,
            mixer_ffn_bias=mixer_ffn_bias, conv_embedding=conv_embedding,
            conv_embedding_kernel_size=conv_embedding_kernel_size)
        act = activations[activation.lower()] if isinstance(activation, str) else activation
        eact = activations[embedding_activation.lower()] if isinstance(embedding_activation, str) \
            else embedding_activation

        self.conv_init = ConvPatcher(embedding_dim=embedding_dim, patch_size=patch_size)
        if not conv_embedding:
            self.embedding_layer = EmbeddingLayer(embedding_dim=embedding_dim, max_num_registers=max_num_registers,
                                                  max_image_size=max_image_size, activation=eact)
        else:
            self.embedding_layer = ConvEmbedding(embedding_dim=embedding_dim, max_num_registers=max_num_registers,
                                                 max_image_size=max_image_size,
                                                 kernel_size=conv_embedding_kernel_size, activation=eact)
        # cosine schedule of the stochastic-depth rate over depth (model.py:82); eval => identity
        st_p = lambda i: cos(arccos(stochastic_depth_p[0]) * (1 - i / num_blocks)
                             + arccos(stochastic_depth_p[1]) * (i / num_blocks))
        self.blocks = nn.ModuleList([
            Block(embedding_dim=embedding_dim, n_head=n_head, activation_func=act, ff_dropout=ffn_dropout,
                  att_dropout=attn_dropout, multiplication_factor=ff_multiplication_factor,
                  conv_kernel_size=conv_kernel_size, conv_activation=act, conv_first=conv_first,
                  conv_block_num=conv_block_num, normalize_qv=normalize_qv, drop_p=st_p(i),
                  mixer_deptwise_bias=mixer_deptwise_bias, mixer_ffn_bias=mixer_ffn_bias, fast_att=fast_att)
            for i in range(num_blocks)])
        self.final_block = FinalBlock(embedding_dim=embedding_dim, n_head=n_head, activation_func=act,
                                      multiplication_factor=ff_multiplication_factor, ff_dropout=ffn_dropout,
                                      att_dropout=attn_dropout, normalize_qv=normalize_qv, drop_p=0.0)
        self.output_head = ClassificationHead(embedding_dim, output_classes, ffn_dropout,
                                              from_register=head_output_from_register,
                                              simple_output=simple_mlp_output, bias=output_head_bias)
        self.__init_weights__()

    def __init_weights__(self):
        # model.py:121-126
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.trunc_normal_(m.weight, std=0.01)

    def engine(self) -> Engine:
        """Packed-weight engine for the current parameters, device and precision (cached)."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("sdpnet_b200.MainModel runs on CUDA only: call .to('cuda') (there is no CPU fallback)")
        return self._packed(lambda: Engine(self._engine_cfg, self.state_dict(), dev, self.precision))

    def forward(self, x: torch.Tensor, num_registers: int = 3, return_raw_outputs: bool = False):
        self._guard(x)
        return self.engine().forward(x, num_registers, return_raw_outputs)
