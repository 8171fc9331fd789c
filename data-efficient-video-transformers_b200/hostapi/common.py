"""Shared plumbing for the host-API modules."""
import warnings

import torch
import torch.nn as nn
from torch.nn import TransformerEncoder, TransformerEncoderLayer

from .. import ops
from ..capi import TvtError
from ..capi import ACT_RELU
from ..functions import EncoderLayerFn, LayerCfg

# inference-only encoder stacks run without LayerNorm kernels when the shapes allow it (run_encoder_folded); TVT_FOLD_LN=0 disables
import os as _os
FOLD_LAYERNORM = _os.environ.get("TVT_FOLD_LN", "1") != "0"


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


def _fold_ok(mode, enc, tokens, training):
    """The LayerNorm-free inference stack applies when nothing is recorded for backward, the mode is bf16, every layer is the
    default post-norm ReLU layer and the layer GEMMs run on the tiles that carry the folded epilogues."""
    if training or torch.is_grad_enabled() or mode.fp32 or tokens.dtype != torch.bfloat16 or len(enc.layers) == 0:
        return False
    n, d = tokens.shape
    for layer in enc.layers:
        if layer.norm_first or layer.activation_relu_or_gelu != 1:
            return False
        ff = layer.linear1.out_features
        if layer.norm1.eps != layer.norm2.eps or d % 32 or ff % 32:
            return False
        if not (ops.ln_fold_supported(n, 3 * d, d) and ops.ln_fold_supported(n, d, d) and ops.ln_fold_supported(n, ff, d)
                and ops.ln_fold_supported(n, d, ff)):
            return False
    return True


def run_encoder_folded(mode, enc, tokens, batch, attn_impl=0, out=None):
    """Inference-only post-norm encoder stack WITHOUT LayerNorm kernels (the frozen teacher, evaluation): a layer's two
    LayerNorm outputs are never materialised.  The GEMM that writes a pre-norm tensor y (out-proj + residual, FFN2 + residual)
    accumulates each row's (sum, sum of squares) in its epilogue; the GEMMs that consume LN(y) as their A operand (FFN1, the next
    layer's packed QKV projection) run on y itself with gamma folded into the weight and the per-row mean / rstd applied in
    their epilogue; the GEMMs that add LN(y) as a residual recompute it from y in theirs (tvt_gemm_args.ln_*).  Only the last
    layer's output is normalised by the LayerNorm kernel, for the consumers outside the stack.
    torch/nn/modules/transformer.py:952-982 is the arithmetic being rearranged.
    ``out=(buf [B, n_blocks * S, d], block)``: the final LayerNorm writes its rows into that block of a wider per-clip buffer
    (ops.layernorm_fwd) - how FusionTransformer lays the memory experts' tokens side by side without a concatenation pass."""
    import math
    n, d = tokens.shape
    L = len(enc.layers)
    slots = 2 * ((d + 255) // 256)     # partial (sum, sum of squares) per half of every 256-column block of the producing GEMM
    stats = torch.empty(2 * L, n, slots, 2, dtype=torch.float32, device=tokens.device)   # rows of y1 / y2 per layer (every slot is written)
    x_real = tokens                # layer input: a materialised tensor (layer 0) ...
    prev = None                    # ... or (y2, stats2, norm2) of the previous layer
    for li, layer in enumerate(enc.layers):
        sa = layer.self_attn
        H = sa.num_heads
        hd, S, ff, eps = d // H, n // batch, layer.linear1.out_features, layer.norm1.eps
        qkv = torch.empty(n, 3 * d, dtype=torch.bfloat16, device=tokens.device)
        if prev is None:
            wh, _ = mode.weight(sa.in_proj_weight)
            ops.gemm(x_real, wh, n, 3 * d, d, bias=sa.in_proj_bias, out_bf16=qkv)
        else:
            y_in, st_in, nrm = prev
            wp, c, bp = mode.folded_weight(sa.in_proj_weight, sa.in_proj_bias, nrm.weight, nrm.bias)
            ops.gemm(y_in, wp, n, 3 * d, d, bias=bp, out_bf16=qkv, ln_in=(st_in, c), ln_dim=d, ln_eps=eps)
        attn, _ = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], batch, H, S, S, hd, 1.0 / math.sqrt(hd), impl=attn_impl)
        # y1 = attn Wo^T + bo + x   (x = the layer input: real, or LN2 of the previous layer recomputed from its y2)
        y1 = torch.empty(n, d, dtype=torch.bfloat16, device=tokens.device)
        st1, st2 = stats[2 * li], stats[2 * li + 1]
        wo, _ = mode.weight(sa.out_proj.weight)
        if prev is None:
            ops.gemm(attn, wo, n, d, d, bias=sa.out_proj.bias, residual=x_real, out_bf16=y1, stats_out=st1)
        else:
            ops.gemm(attn, wo, n, d, d, bias=sa.out_proj.bias, residual=y_in, out_bf16=y1, stats_out=st1,
                     ln_res=(st_in, nrm.weight, nrm.bias), ln_dim=d, ln_eps=eps)
        # h = relu(LN1(y1) W1^T + b1)
        w1p, c1, b1p = mode.folded_weight(layer.linear1.weight, layer.linear1.bias, layer.norm1.weight, layer.norm1.bias)
        h = torch.empty(n, ff, dtype=torch.bfloat16, device=tokens.device)
        ops.gemm(y1, w1p, n, ff, d, bias=b1p, act=ACT_RELU, out_bf16=h, ln_in=(st1, c1), ln_dim=d, ln_eps=eps)
        # y2 = h W2^T + b2 + LN1(y1)
        w2, _ = mode.weight(layer.linear2.weight)
        y2 = torch.empty(n, d, dtype=torch.bfloat16, device=tokens.device)
        ops.gemm(h, w2, n, d, ff, bias=layer.linear2.bias, residual=y1, out_bf16=y2, stats_out=st2,
                 ln_res=(st1, layer.norm1.weight, layer.norm1.bias), ln_dim=d, ln_eps=eps)
        prev = (y2, st2, layer.norm2)
    y2, _, nrm = prev
    res, _, _ = ops.layernorm_fwd(y2, nrm.weight, nrm.bias, nrm.eps, save_stats=False, out=out)
    return res


def blocked_output_ok(mode, enc, tokens, training):
    """True when run_encoder(..., out=...) can write its result straight into a block of a wider buffer."""
    return FOLD_LAYERNORM and enc.norm is None and tokens.shape[1] <= 1024 and _fold_ok(mode, enc, tokens, training)


def run_encoder(mode, enc, tokens, batch, training, attn_impl=0, out=None):
    """Apply every layer of an nn.TransformerEncoder container to batch-major tokens [B*S, d].
    ``out``: see run_encoder_folded (only when blocked_output_ok(...) holds)."""
    if out is not None:
        if not blocked_output_ok(mode, enc, tokens, training):
            raise TvtError("run_encoder: blocked output needs the LayerNorm-folded inference path")
        return run_encoder_folded(mode, enc, tokens, batch, attn_impl, out=out)
    if FOLD_LAYERNORM and _fold_ok(mode, enc, tokens, training):
        tokens = run_encoder_folded(mode, enc, tokens, batch, attn_impl)
        if enc.norm is not None:
            from ..functions import LayerNormFn
            tokens = LayerNormFn.apply(tokens, enc.norm.weight, enc.norm.bias, enc.norm.eps)
        return tokens
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
