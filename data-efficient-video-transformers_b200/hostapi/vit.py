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
        if x.requires_grad:      # an activation from upstream (ViViT): keep the autograd edge (to_act is for raw inputs)
            tok = (x.float() if self.mode.fp32 else x.to(torch.bfloat16)).contiguous().view(b * n, d)
        else:
            tok = to_act(self.mode, x).view(b * n, d)
        for attn, ff in self.layers:
            a, f = attn.fn, ff.fn
            cfg = LayerCfg(self.mode, b, a.heads, self.dropout_p, self.training, "gelu")
            has_out = not isinstance(a.to_out, nn.Identity)
            tok = PreNormLayerFn.apply(cfg, a.dim_head, tok, attn.norm.weight, attn.norm.bias, a.to_qkv.weight,
                                       a.to_out[0].weight if has_out else None, a.to_out[0].bias if has_out else None,
                                       ff.norm.weight, ff.norm.bias, f.net[0].weight, f.net[0].bias, f.net[3].weight, f.net[3].bias)
        return LayerNormFn.apply(tok, self.norm.weight, self.norm.bias, self.norm.eps).view(b, n, d)


class ViViT(nn.Module):
    """Drop-in for src/models/vit.py:79-128 (factorised space / time ViViT): same constructor signature, parameter names
    (``to_patch_embedding.1``, ``pos_embedding``, ``space_token``, ``space_transformer``, ``temporal_token``,
    ``temporal_transformer``, ``mlp_head``) and construction order (so a seeded construction draws the reference's weights).
    The patch projection, both pre-norm transformers and the LayerNorm + Linear head run on the sm_100a kernels; the token
    concat / positional add between them are torch ops (autograd does their backward).  ``precision`` is the extra key."""

    def __init__(self, image_size, patch_size, num_classes, num_frames, dim=192, depth=4, heads=3, pool='cls', in_channels=3,
                 dim_head=64, dropout=0., emb_dropout=0., scale_dim=4, precision="bf16"):
        super().__init__()
        from einops.layers.torch import Rearrange
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        assert image_size % patch_size == 0, 'Image dimensions must be divisible by the patch size.'
        num_patches = (image_size // patch_size) ** 2
        patch_dim = in_channels * patch_size ** 2
        self.mode = ops.Mode(precision)
        self.to_patch_embedding = nn.Sequential(
            Rearrange('b t c (h p1) (w p2) -> b t (h w) (p1 p2 c)', p1=patch_size, p2=patch_size), nn.Linear(patch_dim, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames, num_patches + 1, dim))
        self.space_token = nn.Parameter(torch.randn(1, 1, dim))
        self.space_transformer = Transformer(dim, depth, heads, dim_head, dim * scale_dim, dropout, precision)
        self.temporal_token = nn.Parameter(torch.randn(1, 1, dim))
        self.temporal_transformer = Transformer(dim, depth, heads, dim_head, dim * scale_dim, dropout, precision)
        self.space_transformer.mode = self.temporal_transformer.mode = self.mode
        self.dropout = nn.Dropout(emb_dropout)
        self.pool = pool
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

    def forward(self, x):
        from ..functions import HeadLinearFn, LinearFn
        patches = self.to_patch_embedding[0](x)                                  # b t (h w) (p1 p2 c)
        b, t, n, pd = patches.shape
        proj = self.to_patch_embedding[1]
        tok = LinearFn.apply(self.mode, to_act(self.mode, patches.reshape(b * t * n, pd)), proj.weight, proj.bias).view(b, t, n, -1)
        dim = tok.shape[-1]
        cls_space = self.space_token.to(tok.dtype).expand(b, t, 1, dim)          # repeat '() n d -> b t n d'
        xs = torch.cat((cls_space, tok), dim=2) + self.pos_embedding[:, :, :(n + 1)].to(tok.dtype)
        xs = self.dropout(xs)
        xs = self.space_transformer(xs.reshape(b * t, n + 1, dim))
        xs = xs[:, 0].reshape(b, t, dim)
        cls_temporal = self.temporal_token.to(xs.dtype).expand(b, 1, dim)
        xt = self.temporal_transformer(torch.cat((cls_temporal, xs), dim=1).contiguous())
        pooled = xt.float().mean(dim=1).to(xt.dtype) if self.pool == 'mean' else xt[:, 0]
        h = self.mlp_head
        y = LayerNormFn.apply(pooled.contiguous(), h[0].weight, h[0].bias, h[0].eps)
        return HeadLinearFn.apply(y, h[1].weight, h[1].bias)
