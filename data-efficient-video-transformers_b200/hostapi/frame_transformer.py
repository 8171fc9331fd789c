"""Drop-in for the feature-sequence half of the reference's src/models/frame_transformer.py.

``TransformerBase`` keeps the reference signature (:37-47).  ``FrameStream`` is what FrameTransformer runs
after its CNN backbone has produced one feature vector per scene (:204-210, :176-180): PositionalEncoding
-> TransformerBase -> CLS -> GELU MLP head, plus the training-step losses (:246-282).  The R(2+1)D /
ResNet backbones themselves are out of scope (pretrained weights need the network;
BASELINE.json feeds synthetic *features*).
"""
import torch
import torch.nn as nn

from .. import ops
from ..capi import ACT_GELU, ACT_NONE
from ..compat import LightningModule
from ..functions import DistillLossFn, HeadLinearFn, MlpFn, ReadoutFn
from .common import make_encoder, run_encoder, to_act
from .transformer import PositionalEncoding


class TransformerBase(LightningModule):
    def __init__(self, input_dimension, output_dimension, nhead, nhid, nlayers, dropout, precision="bf16"):
        super().__init__()
        self.transformer = make_encoder(input_dimension, nhead, nhid, dropout, nlayers)
        self.mode = ops.Mode(precision)

    def tokens_forward(self, tokens, batch):
        return run_encoder(self.mode, self.transformer, tokens, batch, self.training)

    def forward(self, x):
        """x (S, B, d) seq-first like the reference -> (S, B, d)."""
        S, B, d = x.shape
        tok = to_act(self.mode, x.transpose(0, 1).reshape(B * S, d))
        return self.tokens_forward(tok, B).view(B, S, d).transpose(0, 1)


class FrameStream(LightningModule):
    """features (B, S, d) -> logits (B, n_classes); ``tokens`` returns the encoded (B, S, d) sequence.
    Attribute names follow FrameTransformer (position_encoder, distil_transformer, img_mlp_head)."""

    def __init__(self, d=896, nhead=2, nhid=512, nlayers=4, dropout=0.5, seq_len=14, n_classes=19, precision="bf16"):
        super().__init__()
        self.mode = ops.Mode(precision)
        self.position_encoder = PositionalEncoding(d, dropout, max_len=seq_len)
        self.distil_transformer = TransformerBase(d, 128, nhead, nhid, nlayers, dropout, precision)
        self.distil_transformer.mode = self.mode
        self.img_mlp_head = nn.Sequential(nn.Linear(d, 512), nn.GELU(), nn.Linear(512, 128), nn.GELU(), nn.Linear(128, n_classes))
        self.running_labels, self.running_logits = [], []

    def tokens(self, feats, inject=None):
        """feats (B, S, d) -> encoded (B, S, d).  ``inject`` (B, d): the other modality's CLS vector appended
        as an extra token before the positional encoding, as FrameTransformer.img_step does in "sum" mode
        (frame_transformer.py:225-226) -> encoded (B, S + 1, d)."""
        B, S, d = feats.shape
        x = to_act(self.mode, feats)
        if inject is not None:
            x = torch.cat((x, inject.to(x.dtype).unsqueeze(1)), dim=1).contiguous()
            S += 1
        tok = self.position_encoder.tokens_forward(x.view(B * S, d), S)
        return self.distil_transformer.tokens_forward(tok, B).view(B, S, d)

    def sum_forward(self, feats, other_cls):
        """FrameTransformer "sum" mode (frame_transformer.py:143-147,225-239): token injection, then
        head(cls + last token)."""
        seq = self.tokens(feats, inject=other_cls)
        return self.head((seq[:, 0] + seq[:, -1]).contiguous())

    def head(self, cls):
        h = self.img_mlp_head
        y = MlpFn.apply(self.mode, (ACT_GELU, ACT_GELU), (0.0, 0.0), cls, h[0].weight, h[2].weight, h[0].bias, h[2].bias)
        return HeadLinearFn.apply(y, h[4].weight, h[4].bias)

    def forward(self, feats):
        B, S, d = feats.shape
        (cls,) = ReadoutFn.apply(B, S, (), self.tokens(feats).view(B * S, d))
        return self.head(cls)

    def training_step(self, batch, batch_idx, teacher_logits=None):
        """frame_transformer.py:246-282: "vid"/"frame" -> BCE; "distil" -> BCE + CE(argmax teacher)."""
        target, feats = batch[0], batch[1]
        logits = self(feats)
        w_ce = 0.0 if teacher_logits is None else 1.0
        losses = DistillLossFn.apply(logits, teacher_logits, target.float(), 1.0, w_ce, 0.0, 1.0)
        if teacher_logits is not None:
            self.log("train/distilloss", losses[2])
            self.log("train/bass_loss", losses[1])
            self.log("train/cossim", losses[4])
        self.log("train/loss", losses[0])
        return losses[0]
