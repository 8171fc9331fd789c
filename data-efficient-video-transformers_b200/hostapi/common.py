"""Shared plumbing for the host-API modules."""
import warnings

import torch
import torch.nn as nn
from torch.nn import TransformerEncoder, TransformerEncoderLayer

from .. import ops
from ..capi import TvtError
from ..functions import EncoderLayerFn, LayerCfg


def make_encoder(d, nhead, nhid, dropout, nlayers, layer=None):
    """Parameter container identical to the reference's (torch's own classes => same init stream and
    state_dict keys, src/models/transformer.py:39-47); its forward is never called on the hot path."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        layer = layer if layer is not None else TransformerEncoderLayer(d, nhead, nhid, dropout)
        return TransformerEncoder(layer, nlayers)


def to_act(mode, x):
    """fp32 (or any float) host-facing tensor -> contiguous activation tensor of the mode's dtype."""
    if not x.is_cuda:
        raise TvtError("input tensor is not on a CUDA device: this path has no CPU implementation")
    if x.dtype == mode.dtype:
        return x.contiguous()
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    if mode.fp32:
        return x
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    ops.split_f32(x, out)
    return out


def run_encoder(mode, enc, tokens, batch, training, attn_impl=0):
    """Apply every layer of an nn.TransformerEncoder container to batch-major tokens [B*S, d]."""
    for layer in enc.layers:
        if layer.norm_first:
            raise TvtError("norm_first encoder layers are handled by hostapi.vit, not run_encoder")
        if layer.activation_relu_or_gelu not in (1, 2):
            raise TvtError("only relu / gelu encoder activations are implemented")
        act = "gelu" if layer.activation_relu_or_gelu == 2 else "relu"
        sa = layer.self_attn
        cfg = LayerCfg(mode, batch, sa.num_heads, layer.dropout.p, training, act, attn_impl)
        tokens = EncoderLayerFn.apply(cfg, tokens, None, sa.in_proj_weight, sa.in_proj_bias, sa.out_proj.weight,
                                      sa.out_proj.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                                      layer.linear2.bias, layer.norm1.weight, layer.norm1.bias, layer.norm2.weight,
                                      layer.norm2.bias)
    if enc.norm is not None:
        from ..functions import LayerNormFn
        tokens = LayerNormFn.apply(tokens, enc.norm.weight, enc.norm.bias, enc.norm.eps)
    return tokens
