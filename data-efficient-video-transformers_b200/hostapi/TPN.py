"""Drop-in for the reference's src/models/TPN.py pyramid parts: ``sum_group`` (:64-72), ``Reasoning``
(:75-112) and the spatial pyramid (:2-40, concat order high, mid, low :58).  The ResNet-34 trunk that
produces the feature maps (TPN.net) is out of scope."""
import torch
import torch.nn as nn

from .. import ops
from ..capi import ACT_RELU
from ..compat import LightningModule
from ..functions import HeadLinearFn, MlpFn, PyramidHeadFn, ReadoutFn
from .common import to_act


def sum_group(x, groups=2):
    """(B, T, d) -> (B, floor(T/groups) * d); remainder frames dropped (TPN.py:64-72).  Forward only."""
    B, T, d = x.shape
    mode = ops.Mode("fp32" if x.dtype == torch.float32 else "bf16")
    xa = to_act(mode, x).view(B * T, d)
    (out,) = ops.pyramid_pool_fwd(xa, B, T, d, [groups], relu=False, skip_cls=False)
    return out


class Reasoning(nn.Module):
    """Temporal pyramid head.  forward(x (B, T, d)) -> (B, num_class) probabilities (mean of per-scale
    sigmoids).  ``forward_tokens`` takes the package's batch-major token matrix and pools the frame
    tokens in place (CLS row skipped)."""

    def __init__(self, num_segments=4, num_frames=5, num_class=15, img_dim=896, max_group=4, start=2, precision="bf16"):
        super().__init__()
        self.num_segments, self.num_frames, self.num_class = num_segments, num_frames, num_class
        self.img_feature_dim, self.num_groups, self.start = img_dim, max_group, start
        self.mode = ops.Mode(precision)
        self.relation = nn.ModuleList()
        self.classifier_scales = nn.ModuleList()
        for scales in range(start, max_group + 1):
            self.relation += [nn.Sequential(
                nn.ReLU(), nn.Linear(img_dim * int(num_segments * num_frames / scales), 512), nn.ReLU(),
                nn.Dropout(p=0.6), nn.Linear(512, 512), nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(512, num_class),
                nn.Sigmoid())]

    @property
    def groups(self):
        return tuple(range(self.start, self.num_groups + 1))

    def scale_logits(self, pooled):
        """pooled: per-scale relu(sum_group) matrices -> z [G, B, C] fp32 (pre-sigmoid)."""
        zs = []
        for rel, xg in zip(self.relation, pooled):
            p1 = rel[3].p if self.training else 0.0
            p2 = rel[6].p if self.training else 0.0
            h = MlpFn.apply(self.mode, (ACT_RELU, ACT_RELU), (p1, p2), xg, rel[1].weight, rel[4].weight, rel[1].bias, rel[4].bias)
            zs.append(HeadLinearFn.apply(h, rel[7].weight, rel[7].bias))
        return torch.stack(zs)

    def forward_tokens(self, tokens, B, S, target=None, skip_cls=True):
        """Returns (prob [B, C], bce loss [1] or None).  Does NOT consume the CLS readout."""
        d = tokens.shape[1]
        if skip_cls:
            outs = ReadoutFn.apply(B, S, self.groups, tokens)[1:]
        else:  # a plain (B, T, d) feature batch: prepend nothing, pool every row
            outs = _PoolAll.apply(B, S, self.groups, tokens)
        z = self.scale_logits(outs)
        if target is None:
            prob, _, _ = ops.pyramid_head(z.detach().contiguous())
            return prob, None
        return PyramidHeadFn.apply(z, target)

    def forward(self, x):
        B, T, d = x.shape
        tok = to_act(self.mode, x).view(B * T, d)
        return self.forward_tokens(tok, B, T, None, skip_cls=False)[0]


class _PoolAll(torch.autograd.Function):
    """sum_group over every row of a (B, T, d) batch (no CLS row to skip)."""

    @staticmethod
    def forward(ctx, B, T, groups, x):
        d = x.shape[1]
        outs = ops.pyramid_pool_fwd(x.contiguous(), B, T, d, groups, relu=True, skip_cls=False)
        ctx.meta = (B, T, d, tuple(groups))
        ctx.save_for_backward(*outs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        B, T, d, groups = ctx.meta
        outs = ctx.saved_tensors
        dx = torch.empty(B * T, d, dtype=outs[0].dtype, device=outs[0].device)
        douts = [torch.zeros_like(o) if t is None else t.contiguous() for t, o in zip(douts, outs)]
        ops.pyramid_pool_bwd(douts, outs, dx, B, T, d, groups, relu=True, skip_cls=False)
        return None, None, None, dx


class SpatialPyramid(nn.Module):
    """Feature_Pyramid_low / Mid / High + concat (TPN.py:2-40,55-58): avg-pool each map to 1x1 (bandwidth
    kernel), 1x1 conv on low and mid (a [C, C] GEMM on the pooled vectors), High pooled only; output
    (frames, 896) in the order (high, mid, low).  Inference path (the maps come from a frozen CNN)."""

    def __init__(self):
        super().__init__()
        self.pyramid_low = nn.ModuleDict({"channels_reduce": nn.Conv2d(128, 128, kernel_size=1)})
        self.pyramid_mid = nn.ModuleDict({"channels_reduce": nn.Conv2d(256, 256, kernel_size=1)})
        self.pyramid_high = nn.ModuleDict({"channels_reduce": nn.Conv2d(512, 512, kernel_size=1)})
        self.mode = ops.Mode("fp32")

    @torch.no_grad()
    def forward(self, low, mid, high):
        frames = low.shape[0]
        out = torch.empty(frames, 896, dtype=torch.float32, device=low.device)
        pooled_mid = torch.empty(frames, 256, dtype=torch.float32, device=low.device)
        pooled_low = torch.empty(frames, 128, dtype=torch.float32, device=low.device)
        ops.spatial_pool(high.contiguous(), out, 0)
        ops.spatial_pool(mid.contiguous(), pooled_mid, 0)
        ops.spatial_pool(low.contiguous(), pooled_low, 0)
        for pooled, conv, off in ((pooled_mid, self.pyramid_mid["channels_reduce"], 512),
                                  (pooled_low, self.pyramid_low["channels_reduce"], 768)):
            Cc = pooled.shape[1]
            w = conv.weight.view(Cc, Cc)
            wh, wl = self.mode.weight(w)
            xp = self.mode.split(pooled)
            ops.gemm(xp[0], wh, frames, Cc, Cc, a_lo=xp[1], b_lo=wl, bias=conv.bias, out_f32=out[:, off:off + Cc])
        return out
