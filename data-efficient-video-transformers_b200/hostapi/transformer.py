"""Drop-in for the reference's src/models/transformer.py (PositionalEncoding, SimpleTransformer).

Same constructor keys (every config.yaml key is accepted as a kwarg and stored in ``hparams``), same
method names (``add_pos_cls``, ``ptn``, ``ptn_shared``, ``shared_step``, ``training_step`` ...), same
parameter names/shapes, same quirks (PE base 1000, one CLS per batch slot, encoders for experts 0 and 1
only).  The literal 2048 / 15 of the reference are ``input_dimension`` / ``n_classes`` here (equal to the
reference at its own values).  Extra optional key: ``precision`` ("bf16" default | "fp32").
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import TransformerEncoderLayer

from .. import ops
from ..compat import LightningModule
from ..functions import EmbedFn, HeadLinearFn, LayerNormFn, PosEncFn, DistillLossFn, ReadoutFn
from .common import make_encoder, run_encoder, to_act


class PositionalEncoding(LightningModule):
    """src/models/transformer.py:10-25.  forward takes seq-first (S, B, d) like the reference."""

    def __init__(self, d_model, dropout=0.1, max_len=4):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(1000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        pe = pe.unsqueeze(0).transpose(0, 1)
        self.register_buffer("pe", pe)

    def tokens_forward(self, tokens, S):
        """Batch-major [B*S, d] entry used inside the package."""
        if S > self.pe.shape[0]:
            raise ValueError(f"sequence length {S} exceeds PositionalEncoding max_len {self.pe.shape[0]}")
        p = self.dropout.p if self.training else 0.0
        return PosEncFn.apply(tokens, self.pe.reshape(-1, self.pe.shape[-1])[:S].contiguous(), S, p)

    def forward(self, x):
        S, B, d = x.shape
        tok = x.transpose(0, 1).reshape(B * S, d)
        return self.tokens_forward(tok.contiguous(), S).view(B, S, d).transpose(0, 1)


class SimpleTransformer(LightningModule):
    def __init__(self, **kwargs):
        super().__init__()
        self.save_hyperparameters()
        hp = self.hparams
        hp.setdefault("dropout", 0.5)
        hp.setdefault("n_classes", 15)
        hp.setdefault("model", "ptn")
        hp.setdefault("precision", "bf16")
        if hp.cls:
            hp.seq_len += 1
        d = hp.input_dimension
        self.mode = ops.Mode(hp.precision)
        self.criterion = nn.BCEWithLogitsLoss()
        self.position_encoder = PositionalEncoding(d, hp.dropout, max_len=hp.seq_len)
        self.encoder_layers0 = TransformerEncoderLayer(d, hp.nhead, hp.nhid, hp.dropout)
        self.transformer_encoder0 = make_encoder(d, hp.nhead, hp.nhid, hp.dropout, hp.nlayers, self.encoder_layers0)
        self.encoder_layers1 = TransformerEncoderLayer(d, hp.nhead, hp.nhid, hp.dropout)
        self.transformer_encoder1 = make_encoder(d, hp.nhead, hp.nhid, hp.dropout, hp.nlayers, self.encoder_layers1)
        self.norm = nn.LayerNorm(d)
        self.running_labels = []
        self.running_logits = []
        self.cls = nn.Parameter(torch.rand(1, hp.batch_size, d))
        self.mlp_head = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, hp.n_classes))
        self.mlp_encoder = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 1024))

    def configure_optimizers(self):
        hp = self.hparams
        return torch.optim.SGD(self.parameters(), lr=hp.learning_rate, momentum=hp.momentum, weight_decay=hp.weight_decay)

    # ---- package-internal batch-major path
    def _embed(self, feat):
        """feat (B, T, d) in the mode's dtype -> tokens [B*S, d]  (add_pos_cls, transformer.py:74-82)."""
        p = self.position_encoder.dropout.p if self.training else 0.0
        return EmbedFn.apply(self.mode, feat, self.cls, self.position_encoder.pe, self.norm.weight, self.norm.bias, p)

    def _head(self, cls_sum):
        x = LayerNormFn.apply(cls_sum, self.mlp_head[0].weight, self.mlp_head[0].bias, self.mlp_head[0].eps)
        return HeadLinearFn.apply(x, self.mlp_head[1].weight, self.mlp_head[1].bias)

    def add_pos_cls(self, data):
        """(B, T, d) -> seq-first (S, B, d), the reference's signature."""
        B, T, d = data.shape
        tok = self._embed(to_act(self.mode, data))
        return tok.view(B, T + 1, d).transpose(0, 1)

    def ptn(self, data):
        """data (BATCH, SEQ, EXPERTS, DIM) -> logits (BATCH, n_classes) fp32   (transformer.py:106-133)."""
        B, T, E, d = data.shape
        toks = []
        for i in range(E):
            tok = self._embed(to_act(self.mode, data[:, :, i, :]))
            if i == 0:
                tok = run_encoder(self.mode, self.transformer_encoder0, tok, B, self.training)
            elif i == 1:
                tok = run_encoder(self.mode, self.transformer_encoder1, tok, B, self.training)
            toks.append(tok)
        (cls_sum,) = ReadoutFn.apply(B, T + 1, (), *toks)
        return self._head(cls_sum)

    def ptn_shared(self, data):
        """transformer.py:84-104 with its intent restored (see oracle.param.SimpleTransformer.ptn_shared)."""
        B, T, E, d = data.shape
        cls_tokens = []
        for i in range(E):
            tok = run_encoder(self.mode, self.transformer_encoder0, self._embed(to_act(self.mode, data[:, :, i, :])), B, self.training)
            cls_tokens.append(tok.view(B, T + 1, d)[:, 0])
        e = torch.stack(cls_tokens, dim=1).contiguous()                      # (B, E, d)
        tok = run_encoder(self.mode, self.transformer_encoder1, self._embed(e), B, self.training)
        (cls,) = ReadoutFn.apply(B, E + 1, (), tok)
        return self._head(cls)

    def shared_step(self, data):
        if self.hparams.model in ("ptn", "ptn_shared"):
            return self.ptn(data)                                            # transformer.py:162-168 (both call ptn)

    def _loss(self, logits, target):
        return DistillLossFn.apply(logits, None, target.reshape(logits.shape), 1.0, 0.0, 0.0, 1.0)[0]

    def training_step(self, batch, batch_idx):
        logits = self.shared_step(batch["experts"])
        loss = self._loss(logits, batch["label"])
        self.log("train/loss", loss, on_step=True, on_epoch=True)
        return loss

    def validation_step(self, batch, batch_idx):
        logits = self.shared_step(batch["experts"])
        target = batch["label"]
        loss = self._loss(logits, target)
        self.log("val/loss", loss, on_step=True, on_epoch=True)
        self.running_logits.append(torch.sigmoid(logits))                    # transformer.py:153-158
        self.running_labels.append(target.int())
        self.running_logits.append(logits)
        return loss

    def format_target(self, target):
        return torch.cat(target, dim=0).squeeze()
