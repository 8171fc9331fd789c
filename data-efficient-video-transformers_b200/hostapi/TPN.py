"""Drop-in for the reference's src/models/TPN.py: ``sum_group`` (:64-72), ``Reasoning`` (:75-112), the spatial
pyramid ``Feature_Pyramid_low / Mid / High`` (:2-40, trainable 1x1 convs, concat order high, mid, low :58) and ``TPN``
(:43-61).  The ResNet-34 trunk that produces the feature maps (TPN.net) is a library plug-in (out of scope as a kernel)."""
import torch
import torch.nn as nn

from .. import ops
from ..capi import ACT_RELU
from ..compat import LightningModule
from ..functions import HeadLinearFn, LinearFn, MlpFn, PyramidHeadFn, ReadoutFn, SpatialPoolFn
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


class _FeaturePyramid(nn.Module):
    """Feature_Pyramid_low / Mid / High (TPN.py:2-40): ``pool_branch`` = AvgPool2d(k) of a k x k map, ``channels_reduce`` =
    a trainable 1x1 conv = a [C, C] Linear on the pooled vector.  The pool is the bandwidth kernel (``SpatialPoolFn``:
    TMA-bulk tile ring, forward and backward), the conv a tensor-core GEMM (``LinearFn``: bias epilogue, wgrad / dgrad /
    bias-gradient backward), so the module trains like the reference's.  forward returns [frames, C, 1, 1] like the
    reference (TPN.forward squeezes it)."""

    C, K, USE_CONV = 0, 0, True

    def __init__(self, precision="fp32"):
        super().__init__()
        self.pool_branch = nn.Sequential(nn.AvgPool2d(kernel_size=self.K))
        self.channels_reduce = nn.Conv2d(self.C, self.C, kernel_size=1)
        self.mode = ops.Mode(precision)

    def pooled(self, x):
        """[frames, C, K, K] -> [frames, C] (after the 1x1 conv when the level has one) in the mode's dtype."""
        if not x.is_cuda:
            raise ops.TvtError("input tensor is not on a CUDA device: this path has no CPU implementation")
        if x.shape[1] != self.C or x.shape[2] != self.K or x.shape[3] != self.K:
            raise ValueError(f"{type(self).__name__}: expected a [*, {self.C}, {self.K}, {self.K}] map, got {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        p = SpatialPoolFn.apply(x)
        if not self.USE_CONV:
            return p
        p = p if self.mode.fp32 else p.to(torch.bfloat16)
        w = self.channels_reduce.weight
        return LinearFn.apply(self.mode, p, w.view(self.C, self.C), self.channels_reduce.bias).float()

    def forward(self, x):
        return self.pooled(x).view(-1, self.C, 1, 1)


class Feature_Pyramid_Mid(_FeaturePyramid):
    C, K = 256, 14


class Feature_Pyramid_High(_FeaturePyramid):
    """TPN.py:16-26: the High level constructs its conv but returns the pooled tensor without applying it."""
    C, K, USE_CONV = 512, 7, False


class Feature_Pyramid_low(_FeaturePyramid):
    C, K = 128, 28


class ResNetMaps(nn.Module):
    """The (layer2, layer3, layer4) feature maps of a torchvision ResNet — what the reference's
    ``custom_resnet.ResNet.forward`` returns (custom_resnet.py:138-153).  Library CNN, out of scope as a kernel."""

    def __init__(self, arch="resnet34", pretrained=False):
        super().__init__()
        import torchvision.models as models
        net = getattr(models, arch)(weights="DEFAULT" if pretrained else None)
        for name in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4", "avgpool", "fc"):
            setattr(self, name, getattr(net, name))

    def forward(self, x):
        x = self.layer1(self.maxpool(self.relu(self.bn1(self.conv1(x)))))
        x2 = self.layer2(x)
        x3 = self.layer3(x2)
        return x2, x3, self.layer4(x3)


class TPN(LightningModule):
    """TPN.py:43-61: CNN trunk -> spatial pyramid (low / mid / high, concat order high, mid, low -> 896) -> ``Reasoning``
    over the frames.  ``net`` is a plug-in returning the three maps (default: torchvision's ResNet-34, no download)."""

    def __init__(self, net=None, precision="fp32", pretrained=False):
        super().__init__()
        self.net = net if net is not None else ResNetMaps("resnet34", pretrained)
        self.pyramid_low = Feature_Pyramid_low(precision)
        self.pyramid_mid = Feature_Pyramid_Mid(precision)
        self.pyramid_high = Feature_Pyramid_High(precision)
        self.reason = Reasoning(precision=precision)

    def frame_features(self, low, mid, high):
        """The three maps of N frames -> [N, 896] in the reference's concat order (TPN.py:55-58)."""
        return torch.cat((self.pyramid_high.pooled(high), self.pyramid_mid.pooled(mid), self.pyramid_low.pooled(low)), dim=-1)

    def forward(self, x):
        low, mid, high = self.net(x)
        cnn_out = self.frame_features(low, mid, high).unsqueeze(0)        # TPN.py:58: one clip of N frames
        return self.reason(cnn_out)


class SpatialPyramid(nn.Module):
    """The three pyramid levels + concat as one module (``TPN`` without its CNN and Reasoning): maps -> (frames, 896) in
    the order (high, mid, low).  Submodule / parameter names are TPN's (pyramid_low.channels_reduce.weight ...)."""

    def __init__(self, precision="fp32"):
        super().__init__()
        self.pyramid_low = Feature_Pyramid_low(precision)
        self.pyramid_mid = Feature_Pyramid_Mid(precision)
        self.pyramid_high = Feature_Pyramid_High(precision)

    def forward(self, low, mid, high):
        return torch.cat((self.pyramid_high.pooled(high), self.pyramid_mid.pooled(mid), self.pyramid_low.pooled(low)), dim=-1)
