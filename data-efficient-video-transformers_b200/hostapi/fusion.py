"""The BASELINE.json model family (configs 2-5) assembled from the reference's parts plus the two
north-star extensions the reference only names: queries-vs-keys cross-modal attention
(TransformerDecoderLayer is imported at src/models/transformer.py:6 and never used) and a KL term in the
distillation loss (src/models/frame_transformer.py:250-252 has hard-label CE only).  The CPU oracle of
every class here is the class of the same name in oracle/param.py; state_dict keys are identical.
"""
import torch
import torch.nn as nn

from .. import ops
from ..compat import LightningModule
from ..functions import (DistillLossFn, EmbedFn, EncoderLayerFn, HeadLinearFn, LayerCfg, LayerNormFn, LinearFn,
                         PyramidHeadFn, ReadoutFn)
from .common import blocked_output_ok, make_encoder, run_encoder, to_act
from .TPN import Reasoning
from .transformer import PositionalEncoding


class CrossModalBlock(nn.Module):
    """x = LN1(x + Drop(MHA(q=x, k=mem, v=mem)));  x = LN2(x + Drop(W2 Drop(relu(W1 x))))  (post-norm)."""

    def __init__(self, d, nhead, nhid, dropout):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d, nhid)
        self.linear2 = nn.Linear(nhid, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.dropout = nn.Dropout(dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def tokens_forward(self, mode, x, mem, batch):
        a = self.multihead_attn
        cfg = LayerCfg(mode, batch, a.num_heads, self.dropout.p, self.training, "relu")
        return EncoderLayerFn.apply(cfg, x, mem, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                                    self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                                    self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias)


class ExpertStream(nn.Module):
    """Linear(D_e, d) -> per-batch-slot CLS -> PE -> LN -> L-layer encoder; (B, T, D_e) -> tokens [B*S, d]."""

    def __init__(self, in_dim, d, nhead, nhid, nlayers, dropout, batch_size, frames):
        super().__init__()
        self.expert_encoder = nn.Linear(in_dim, d)
        self.position_encoder = PositionalEncoding(d, dropout, max_len=frames + 1)
        self.norm = nn.LayerNorm(d)
        self.cls = nn.Parameter(torch.rand(1, batch_size, d))
        self.transformer_encoder = make_encoder(d, nhead, nhid, dropout, nlayers)

    def tokens_forward(self, mode, x, out=None):
        """``out=(buf [B, n_blocks * S, d], block)`` (inference only): the tokens are ALSO laid into that block of a wider per-clip
        buffer - by the last LayerNorm itself on the LayerNorm-folded path, else by one strided copy."""
        B, T, D = x.shape
        xa = to_act(mode, x).view(B * T, D)
        feat = LinearFn.apply(mode, xa, self.expert_encoder.weight, self.expert_encoder.bias).view(B, T, -1)
        p = self.position_encoder.dropout.p if self.training else 0.0
        tok = EmbedFn.apply(mode, feat, self.cls, self.position_encoder.pe, self.norm.weight, self.norm.bias, p)
        if out is None:
            return run_encoder(mode, self.transformer_encoder, tok, B, self.training)
        if blocked_output_ok(mode, self.transformer_encoder, tok, self.training):
            return run_encoder(mode, self.transformer_encoder, tok, B, self.training, out=out)
        res = run_encoder(mode, self.transformer_encoder, tok, B, self.training)
        buf, blk = out
        S = T + 1
        buf[:, blk * S:(blk + 1) * S].copy_(res.view(B, S, -1))
        return res


class FusionTransformer(LightningModule):
    """experts -> ExpertStream each -> fusion ("sum" of CLS tokens | "cross" attention block with queries
    from expert 0 and keys/values from the other experts) -> LN + Linear head -> logits; optional temporal
    pyramid (Reasoning) over the frame tokens.  forward(list of (B, T, D_e)) -> (logits, pyramid probs)."""

    def __init__(self, in_dims, d=512, nhead=8, nhid=2048, nlayers=4, dropout=0.0, batch_size=8, frames=16,
                 n_classes=15, fusion="sum", pyramid=False, max_group=4, precision="bf16"):
        super().__init__()
        in_dims = tuple(in_dims)
        self.save_hyperparameters()
        self.mode = ops.Mode(precision)
        self.streams = nn.ModuleList([ExpertStream(D, d, nhead, nhid, nlayers, dropout, batch_size, frames) for D in in_dims])
        self.fusion = fusion
        if fusion == "cross":
            self.cross = CrossModalBlock(d, nhead, nhid, dropout)
        self.mlp_head = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, n_classes))
        self.reason = Reasoning(1, frames, n_classes, d, max_group, 2, precision) if pyramid else None
        if self.reason is not None:
            self.reason.mode = self.mode

    def forward(self, experts, target=None):
        """Returns (logits [B, C] fp32, pyramid probs [B, C] or None, pyramid BCE loss [1] or None)."""
        m = self.mode
        B, T = experts[0].shape[0], experts[0].shape[1]
        S = T + 1
        E = len(self.streams)
        mem = None
        if self.fusion == "cross" and E > 2 and not torch.is_grad_enabled():
            # inference: every memory expert's last LayerNorm writes its tokens straight into its block of the per-clip memory
            # [B, (E - 1) * S, d] (transformer.py:110-121 concatenates them): no torch.cat pass (101 MB per C5 step)
            d = self.streams[0].expert_encoder.out_features
            buf = torch.empty(B, (E - 1) * S, d, dtype=m.dtype, device=experts[0].device)
            toks = [self.streams[0].tokens_forward(m, experts[0])]
            toks += [s.tokens_forward(m, x, out=(buf, e)) for e, (s, x) in enumerate(zip(self.streams[1:], experts[1:]))]
            mem = buf.view(-1, d)
        else:
            toks = [s.tokens_forward(m, x) for s, x in zip(self.streams, experts)]
        if self.fusion == "cross" and len(toks) > 1:
            if mem is not None:
                pass
            elif len(toks) > 2:   # memory = other experts' tokens concatenated along the sequence, per clip
                d = toks[0].shape[1]
                mem = torch.cat([t.view(B, S, d) for t in toks[1:]], dim=1).reshape(-1, d)
            else:
                mem = toks[1]
            seq = self.cross.tokens_forward(m, toks[0], mem, B)
            readout = [seq]
        else:
            seq = toks[0]
            readout = toks
        groups = self.reason.groups if self.reason is not None else ()
        if groups:
            # CLS from every readout stream, pooled frames from the fused sequence (readout[0] is seq)
            outs = ReadoutFn.apply(B, S, groups, *readout)
            cls, pooled = outs[0], outs[1:]
        else:
            (cls,) = ReadoutFn.apply(B, S, (), *readout)
            pooled = None
        x = LayerNormFn.apply(cls, self.mlp_head[0].weight, self.mlp_head[0].bias, self.mlp_head[0].eps)
        logits = HeadLinearFn.apply(x, self.mlp_head[1].weight, self.mlp_head[1].bias)
        prob = ploss = None
        if pooled is not None:
            z = self.reason.scale_logits(pooled)
            if target is not None:
                prob, ploss = PyramidHeadFn.apply(z, target)
            else:
                prob, _, _ = ops.pyramid_head(z.detach().contiguous())
        return logits, prob, ploss


class DistillationTrainer(LightningModule):
    """Frozen multi-modal teacher -> student distillation step (BASELINE.json configs 4-5):
    teacher forward under no_grad, student forward + backward,
    loss = BCE(s, y) + CE(s, argmax t) + alpha * T^2 * KL + BCE(pyramid probs, y)."""

    def __init__(self, teacher, student, temperature=2.0, alpha=1.0, student_experts=(0,)):
        super().__init__()
        self.teacher, self.student = teacher, student
        self.temperature, self.alpha = float(temperature), float(alpha)
        self.student_experts = tuple(student_experts)
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self.teacher.eval()

    def train(self, mode=True):
        super().train(mode)
        self.teacher.eval()
        return self

    def training_step(self, batch, batch_idx=0):
        experts, target = batch["experts"], batch["label"]
        with torch.no_grad():
            t_logits, _, _ = self.teacher(experts)
        s_logits, _, ploss = self.student([experts[i] for i in self.student_experts], target)
        losses = DistillLossFn.apply(s_logits, t_logits, target.float(), 1.0, 1.0, self.alpha, self.temperature)
        loss = losses[0] if ploss is None else losses[0] + ploss[0]
        self.log("train/loss", loss)
        self.log("train/cossim", losses[4])
        return loss

    def configure_optimizers(self):
        return torch.optim.AdamW([p for p in self.student.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01)
