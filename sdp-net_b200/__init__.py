"""sdpnet_b200 -- B200-native (sm_100a) forward engine for SdP-Net behind the reference's module API.

    import sdpnet_b200 as sdp
    model = sdp.MainModel.from_dict(**model_config).eval().to("cuda")
    model.load_state_dict(reference_state_dict)           # strict, reference key layout
    logits = model(images, num_registers=4)               # one sdp_forward C-ABI call

Compute happens only in `lib/libsdpnet_b200.so` (hand-written CUDA, C-ABI in include/sdpnet_b200.h).
"""
from . import _lib, checkpoint, engine, evaluate, ops, preprocess, shard
from .engine import Engine
from .layers import (Block, ClassificationHead, ConvEmbedding, ConvMixer, ConvPatcher, EmbeddingLayer,
                     EncoderLayer, FinalBlock, LayerNorm, StochasticDepth)
from .model import MainModel, SdPModel, activations
from .preprocess import ValTransforms, val_transforms

__all__ = ["MainModel", "SdPModel", "Engine", "Block", "ClassificationHead", "ConvEmbedding", "ConvMixer",
           "ConvPatcher", "EmbeddingLayer", "EncoderLayer", "FinalBlock", "LayerNorm", "StochasticDepth",
           "activations", "ops", "engine", "checkpoint", "evaluate", "preprocess", "shard", "ValTransforms",
           "val_transforms"]
