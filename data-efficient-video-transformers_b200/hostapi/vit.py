"""Drop-in for the transformer of the reference's src/models/vit.py (PreNorm / FeedForward / Attention /
Transformer, :8-75): pre-norm blocks, bias-free packed qkv, scale dim_head**-0.5, exact-erf GELU MLP, final
LayerNorm; batch-first (b, n, dim) like the reference.  Module layout (ModuleList of [PreNorm(attn),
PreNorm(ff)]) reproduces the reference's state_dict keys.  Dropout inside the attention probabilities does
not exist in the reference (vit.py:53 has no dropout on `attn`) and is not applied here either."""
import torch
import torch.nn as nn

from .. import ops
from ..functions import LayerCfg, LayerNormFn, PreNormLayerFn
from .common import to_act


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden_dim, dim),
                                 nn.Dropout(dropout))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0, precision="bf16"):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.norm = nn.LayerNorm(dim)
        self.dropout_p = dropout
        self.mode = ops.Mode(precision)
        for _ in range(depth):
            self.layers.append(nn.ModuleList([PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                                              PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout))]))

    def forward(self, x):
        b, n, d = x.shape
        tok = to_act(self.mode, x).view(b * n, d)
        for attn, ff in self.layers:
            a, f = attn.fn, ff.fn
            cfg = LayerCfg(self.mode, b, a.heads, self.dropout_p, self.training, "gelu")
            has_out = not isinstance(a.to_out, nn.Identity)
            tok = PreNormLayerFn.apply(cfg, a.dim_head, tok, attn.norm.weight, attn.norm.bias, a.to_qkv.weight,
                                       a.to_out[0].weight if has_out else None, a.to_out[0].bias if has_out else None,
                                       ff.norm.weight, ff.norm.bias, f.net[0].weight, f.net[0].bias, f.net[3].weight, f.net[3].bias)
        return LayerNormFn.apply(tok, self.norm.weight, self.norm.bias, self.norm.eps).view(b, n, d)
