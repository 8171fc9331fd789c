"""Parameterised CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Each class follows the reference file:line given in its docstring (paths relative to /root/reference).
The only deliberate difference from the reference is that widths hard-coded there (2048 in
src/models/transformer.py, 896 in src/models/frame_transformer.py and src/models/TPN.py) are arguments
here; at the reference's own widths ``make_golden.py`` proves the two produce identical numbers.
Attribute names match the reference so ``state_dict`` keys are interchangeable.
"""
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import TransformerEncoder, TransformerEncoderLayer


def _stack(layer, nlayers):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "enable_nested_tensor is True, but ..." (seq-first layers)
        return TransformerEncoder(layer, nlayers)


def _encoder(d, nhead, nhid, dropout, nlayers):
    # src/models/transformer.py:39-47 — default TransformerEncoderLayer: ReLU, post-norm, seq-first, eps 1e-5.
    return _stack(TransformerEncoderLayer(d, nhead, nhid, dropout), nlayers)


class PositionalEncoding(nn.Module):
    """src/models/transformer.py:10-25 (dup src/models/frame_transformer.py:19-34): sinusoid with base
    1000.0 (not 10000), buffer ``pe`` of shape (max_len, 1, d), ``x + pe[:S]`` then dropout."""

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

    def forward(self, x):
        x = x + self.pe[: x.size(0), :]
        return self.dropout(x)


class SimpleTransformer(nn.Module):
    """src/models/transformer.py:28-133 with the literal 2048 replaced by ``input_dimension`` and the
    literal 15 by ``n_classes``.  Construction order of parameters is the reference's, so seeding torch
    and constructing this class draws the same initial weights as constructing the reference class.
    ``seq_len`` is the number of scenes WITHOUT the CLS token (the reference adds 1 when ``cls``)."""

    def __init__(self, input_dimension=2048, nhead=8, nhid=2048, nlayers=8, dropout=0.5, batch_size=2,
                 seq_len=13, cls=1, n_classes=15, **_unused):
        super().__init__()
        d = input_dimension
        self.d = d
        self.seq_len = seq_len + (1 if cls else 0)                      # transformer.py:33-34
        self.criterion = nn.BCEWithLogitsLoss()                         # :35
        self.position_encoder = PositionalEncoding(d, dropout, max_len=self.seq_len)  # :36-37
        # :39-47 — the template layers stay registered (extra state_dict keys encoder_layers{0,1}.*)
        self.encoder_layers0 = TransformerEncoderLayer(d, nhead, nhid, dropout)
        self.transformer_encoder0 = _stack(self.encoder_layers0, nlayers)
        self.encoder_layers1 = TransformerEncoderLayer(d, nhead, nhid, dropout)
        self.transformer_encoder1 = _stack(self.encoder_layers1, nlayers)
        self.norm = nn.LayerNorm(d)                                     # :49
        self.cls = nn.Parameter(torch.rand(1, batch_size, d))           # :52-53 one CLS per batch slot
        self.mlp_head = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, n_classes))      # :54
        self.mlp_encoder = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 1024))        # :55-56 (unused)

    def add_pos_cls(self, data):
        """transformer.py:74-82: (B,T,d) -> (S,B,d): prepend CLS, add PE (+dropout), LayerNorm over d."""
        data = data.transpose(0, 1)
        data = torch.cat((self.cls, data))
        data = self.position_encoder(data)
        data = self.norm(data.transpose(0, 1)).transpose(0, 1)
        return data

    def ptn(self, data):
        """transformer.py:106-133: per-expert encoder (experts 0 and 1 only), CLS gather, sum, head."""
        expert_array = []
        for i, expert in enumerate(data.permute(2, 0, 1, 3)):   # 'b s e d -> e b s d'
            e = self.add_pos_cls(expert)
            if i == 0:
                e = self.transformer_encoder0(e)
            elif i == 1:
                e = self.transformer_encoder1(e)
            expert_array.append(e.transpose(0, 1)[:, 0, :])
        ptn_out = torch.stack(expert_array).transpose(0, 1).sum(dim=1)
        return self.mlp_head(ptn_out)

    def ptn_shared(self, data):
        """transformer.py:84-104 with its intent restored: the broken ``self(e)`` call (forward needs the
        non-existent ``expert_encoder``) is read as ``transformer_encoder0`` for the expert level and
        ``transformer_encoder1`` for the level over expert CLS tokens.  Requires E + 1 <= seq_len."""
        cls_tokens = []
        for expert in data.permute(2, 0, 1, 3):
            e = self.transformer_encoder0(self.add_pos_cls(expert))
            cls_tokens.append(e.transpose(0, 1)[:, 0])
        e = torch.stack(cls_tokens).transpose(0, 1)             # (B, E, d)
        e = self.transformer_encoder1(self.add_pos_cls(e))
        return self.mlp_head(e.transpose(0, 1)[:, 0])

    def loss(self, data, target):
        """transformer.py:135-144."""
        return self.criterion(self.ptn(data), target)


class TransformerBase(nn.Module):
    """src/models/frame_transformer.py:37-47."""

    def __init__(self, input_dimension, output_dimension, nhead, nhid, nlayers, dropout):
        super().__init__()
        self.transformer = _encoder(input_dimension, nhead, nhid, dropout, nlayers)

    def forward(self, x):
        return self.transformer(x)


class FrameStream(nn.Module):
    """The feature-sequence half of FrameTransformer.vid_step / forward("vid")
    (src/models/frame_transformer.py:204-210, :176-180, :99, :106): backbone features (B, S, d) ->
    permute to (S, B, d) -> PositionalEncoding(d, p, max_len=S) -> TransformerBase(d, ., nhead, nhid, L, p)
    -> CLS = token 0 -> GELU MLP head d -> 512 -> 128 -> C.  The CNN backbone that produces the features
    is out of scope."""

    def __init__(self, d=896, nhead=2, nhid=512, nlayers=4, dropout=0.5, seq_len=14, n_classes=19):
        super().__init__()
        self.position_encoder = PositionalEncoding(d, dropout, max_len=seq_len)
        self.distil_transformer = TransformerBase(d, 128, nhead, nhid, nlayers, dropout)
        self.img_mlp_head = nn.Sequential(nn.Linear(d, 512), nn.GELU(), nn.Linear(512, 128), nn.GELU(),
                                          nn.Linear(128, n_classes))

    def tokens(self, feats, inject=None):
        data = feats.permute(1, 0, 2)
        if inject is not None:                      # frame_transformer.py:225-226 ("sum": cat the other modality's CLS)
            data = torch.cat((data, inject.unsqueeze(0)))
        data = self.position_encoder(data)
        return self.distil_transformer(data).permute(1, 0, 2)

    def sum_forward(self, feats, other_cls):
        """frame_transformer.py:143-147,237-239: (cls, last token) -> head(cls + last)."""
        seq = self.tokens(feats, inject=other_cls)
        return self.img_mlp_head(seq[:, 0] + seq[:, -1])

    def forward(self, feats):
        return self.img_mlp_head(self.tokens(feats)[:, 0])


def sum_group(x, groups=2):
    """src/models/TPN.py:64-72: sum consecutive groups of frames, concatenate on the feature axis,
    remainder frames dropped (int(pics / groups))."""
    batch, pics, vector = x.size()
    out = [x[:, groups * g: groups * (g + 1), :].sum(dim=1) for g in range(int(pics / groups))]
    return torch.cat(out, dim=1)


class Reasoning(nn.Module):
    """src/models/TPN.py:75-112: per-scale ReLU -> Linear -> ReLU -> Dropout(.6) -> Linear -> ReLU ->
    Dropout(.5) -> Linear -> Sigmoid, averaged over scales start..max_group."""

    def __init__(self, num_segments=4, num_frames=5, num_class=15, img_dim=896, max_group=4, start=2):
        super().__init__()
        self.num_groups, self.start = max_group, start
        self.relation = nn.ModuleList()
        self.classifier_scales = nn.ModuleList()
        for scales in range(start, max_group + 1):
            self.relation += [nn.Sequential(
                nn.ReLU(), nn.Linear(img_dim * int(num_segments * num_frames / scales), 512), nn.ReLU(),
                nn.Dropout(p=0.6), nn.Linear(512, 512), nn.ReLU(), nn.Dropout(p=0.5),
                nn.Linear(512, num_class), nn.Sigmoid())]

    def forward(self, x):
        prediction = 0
        for g in range(self.start, self.num_groups + 1):
            prediction = prediction + self.relation[g - self.start](sum_group(x, groups=g))
        return prediction / (self.num_groups - self.start + 1)


class SpatialPyramid(nn.Module):
    """src/models/TPN.py:2-40,55-58: AvgPool to 1x1 on (128,28,28) / (256,14,14) / (512,7,7) maps, 1x1 conv
    on low and mid only (the High branch returns the pooled tensor without its conv, :24-26), concat in
    the order (high, mid, low) -> 896 per frame.  Submodule names match TPN's."""

    def __init__(self):
        super().__init__()
        self.pyramid_low = nn.ModuleDict({"channels_reduce": nn.Conv2d(128, 128, kernel_size=1)})
        self.pyramid_mid = nn.ModuleDict({"channels_reduce": nn.Conv2d(256, 256, kernel_size=1)})
        self.pyramid_high = nn.ModuleDict({"channels_reduce": nn.Conv2d(512, 512, kernel_size=1)})

    def forward(self, low, mid, high):
        low_0 = self.pyramid_low["channels_reduce"](F.avg_pool2d(low, 28)).flatten(1)
        mid_0 = self.pyramid_mid["channels_reduce"](F.avg_pool2d(mid, 14)).flatten(1)
        high_0 = F.avg_pool2d(high, 7).flatten(1)
        return torch.cat((high_0, mid_0, low_0), dim=-1)


# ---------------------------------------------------------------------------------------------- vit.py
class VitAttention(nn.Module):
    """src/models/vit.py:30-58: bias-free qkv, scale dim_head**-0.5, explicit softmax(QK^T)V, out proj."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        project_out = not (heads == 1 and dim_head == dim)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if project_out else nn.Identity()

    def forward(self, x):
        b, n, _ = x.shape
        q, k, v = (t.reshape(b, n, self.heads, -1).transpose(1, 2) for t in self.to_qkv(x).chunk(3, dim=-1))
        attn = (torch.einsum("bhid,bhjd->bhij", q, k) * self.scale).softmax(dim=-1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class VitTransformer(nn.Module):
    """src/models/vit.py:8-28,60-75: pre-norm blocks x = attn(LN(x)) + x; x = ff(LN(x)) + x; final LN.
    Layer layout (ModuleList of [PreNorm(attn), PreNorm(ff)]) mirrors the reference's state_dict."""

    class _PreNorm(nn.Module):
        def __init__(self, dim, fn):
            super().__init__()
            self.norm = nn.LayerNorm(dim)
            self.fn = fn

        def forward(self, x):
            return self.fn(self.norm(x))

    class _FeedForward(nn.Module):
        def __init__(self, dim, hidden, dropout=0.0):
            super().__init__()
            self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(dropout),
                                     nn.Linear(hidden, dim), nn.Dropout(dropout))

        def forward(self, x):
            return self.net(x)

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.norm = nn.LayerNorm(dim)
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                self._PreNorm(dim, VitAttention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                self._PreNorm(dim, self._FeedForward(dim, mlp_dim, dropout=dropout))]))

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return self.norm(x)


# ------------------------------------------------------------------------- north-star extensions
class CrossModalBlock(nn.Module):
    """Queries-vs-keys cross-modal attention block.  NOT in the reference (it imports
    TransformerDecoderLayer at src/models/transformer.py:6 and never uses it); torch-pinned: this is
    TransformerDecoderLayer's cross-attention + feed-forward halves (torch/nn/modules/transformer.py
    _mha_block / _ff_block, post-norm, ReLU), i.e.
        x = LN1(x + Drop(MHA(q=x, k=mem, v=mem)));  x = LN2(x + Drop(W2 Drop(relu(W1 x))))"""

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

    def forward(self, x, mem):
        x = self.norm1(x + self.dropout1(self.multihead_attn(x, mem, mem, need_weights=False)[0]))
        return self.norm2(x + self.dropout2(self.linear2(self.dropout(F.relu(self.linear1(x))))))


class ExpertStream(nn.Module):
    """One expert's path of SimpleTransformer.ptn (src/models/transformer.py:74-82,111-118) preceded by the
    input projection the reference names but never defines (``expert_encoder``, transformer.py:67):
    Linear(D_e, d) -> prepend per-batch-slot CLS -> PE -> LN -> L-layer encoder.  Returns (S, B, d)."""

    def __init__(self, in_dim, d, nhead, nhid, nlayers, dropout, batch_size, frames):
        super().__init__()
        self.expert_encoder = nn.Linear(in_dim, d)
        self.position_encoder = PositionalEncoding(d, dropout, max_len=frames + 1)
        self.norm = nn.LayerNorm(d)
        self.cls = nn.Parameter(torch.rand(1, batch_size, d))
        self.transformer_encoder = _encoder(d, nhead, nhid, dropout, nlayers)

    def forward(self, x):
        data = self.expert_encoder(x).transpose(0, 1)
        data = self.position_encoder(torch.cat((self.cls, data)))
        data = self.norm(data.transpose(0, 1)).transpose(0, 1)
        return self.transformer_encoder(data)


class FusionTransformer(nn.Module):
    """The BASELINE.json model family (configs 2-5) assembled from the restated reference parts:
      experts -> ExpertStream each -> fusion -> CLS -> mlp_head (LN + Linear, transformer.py:54) -> logits
      fusion = "sum":   sum of expert CLS tokens (transformer.py:127-130)
      fusion = "cross": CrossModalBlock(q = expert-0 tokens, k = v = other experts' tokens concatenated)
      pyramid: Reasoning over the frame tokens (CLS excluded) of the fused / expert-0 sequence
               (TPN.py:106-112 applied to transformer tokens instead of CNN frame features)
    forward returns (logits [B,C], pyramid probabilities [B,C] or None)."""

    def __init__(self, in_dims, d=512, nhead=8, nhid=2048, nlayers=4, dropout=0.0, batch_size=8, frames=16,
                 n_classes=15, fusion="sum", pyramid=False, max_group=4):
        super().__init__()
        self.streams = nn.ModuleList(
            [ExpertStream(D, d, nhead, nhid, nlayers, dropout, batch_size, frames) for D in in_dims])
        self.fusion = fusion
        if fusion == "cross":
            self.cross = CrossModalBlock(d, nhead, nhid, dropout)
        self.mlp_head = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, n_classes))
        self.reason = Reasoning(1, frames, n_classes, d, max_group, 2) if pyramid else None

    def forward(self, experts):
        toks = [s(x) for s, x in zip(self.streams, experts)]       # each (S, B, d)
        if self.fusion == "cross" and len(toks) > 1:
            seq = self.cross(toks[0], torch.cat(toks[1:], dim=0))
            cls = seq[0]
        else:
            seq = toks[0]
            cls = torch.stack([t[0] for t in toks]).sum(dim=0)
        logits = self.mlp_head(cls)
        pyr = self.reason(seq[1:].transpose(0, 1)) if self.reason is not None else None
        return logits, pyr


def distill_loss(student, teacher, target, temperature=0.0, alpha=0.0, pyramid=None):
    """src/models/frame_transformer.py:250-257: BCEWithLogits(student, target) + CE(student,
    argmax(teacher)); monitor cosine(student, teacher)[0].  North-star extension (torch-pinned): when
    alpha > 0 add alpha * T^2 * kl_div(log_softmax(s/T), softmax(t/T), 'batchmean'); when ``pyramid``
    (probabilities from Reasoning) is given add BCE(pyramid, target).  Returns (loss, parts dict)."""
    base = F.binary_cross_entropy_with_logits(student, target)
    hard = F.cross_entropy(student, torch.argmax(teacher, dim=-1))
    loss = base + hard
    parts = {"base": base, "distil": hard, "cos": F.cosine_similarity(student, teacher, dim=1)[0]}
    if alpha > 0:
        T = temperature
        kl = F.kl_div(F.log_softmax(student / T, dim=-1), F.softmax(teacher / T, dim=-1), reduction="batchmean") * (T * T)
        parts["kl"] = kl
        loss = loss + alpha * kl
    if pyramid is not None:
        pb = F.binary_cross_entropy(pyramid, target)
        parts["pyramid"] = pb
        loss = loss + pb
    return loss, parts


def eval_readout(logits, target, thresholds=(0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8)):
    """What the reference's evaluation path produces per batch: ``F.sigmoid(data)`` and ``target.int()``
    (src/models/transformer.py:150-153), ``(running_logits > threshold).to(int)`` for the callback's thresholds
    (src/callbacks/callbacks.py:37-40), plus the arg-max class used by the top-1 agreement criterion."""
    logits = logits.float()
    probs = torch.sigmoid(logits)
    labels = None if target is None else target.reshape(logits.shape).int()
    preds = [(probs > t).int() for t in thresholds]
    return probs, labels, preds, torch.argmax(logits, dim=-1)


class GatedEmbeddingUnit(nn.Module):
    """src/models/collabgating.py:59-70: Linear then L2 normalisation along dim 1 (its context gate is commented out)."""

    def __init__(self, input_dimension, output_dimension, use_bn=False):
        super().__init__()
        self.fc = nn.Linear(input_dimension, output_dimension)

    def forward(self, x):
        return F.normalize(self.fc(x))


class CollaborativeGating(nn.Module):
    """Restatement of src/models/collabgating.py:3-56 on stacked rows instead of nested Python lists.

    Reference input: list (batch) of list (scenes) of list (experts) of ``[1, D_e]`` tensors; here ``xs`` is a list over
    experts of ``[B, S, D_e]`` tensors and the result is the reference's ``[B, S, 1024]``.  Per (clip, scene), with
    P = ``projection`` (ONE shared Linear(2048, 2048), :9) and the work list L = [x_0 .. x_{E-1}]:
      for i in range(E):  cur = L.pop(0); c = P(pad(cur))                                   (:28-33)
                          t = sum_{o in L} (c + P(pad(o)))                                   (:35-43)
                          out_i = GLU(cat(c, c + P(t))) = c * sigmoid(c + P(t))              (:45-47, :83-85)
                          L.append(c)      <- the PROJECTED vector: later iterations project it again  (:49)
      result = normalize(geu.fc(sum_i out_i))                                                (:50-53, :66-69)
    ``pad`` is nearest-neighbour ``F.interpolate`` to 2048 for any expert whose width differs (:12-16).  E >= 2 (the
    reference stacks an empty list for a single expert)."""

    def __init__(self):
        super().__init__()
        self.proj_input = 2048
        self.proj_embedding_size = 2048
        self.projection = nn.Linear(self.proj_input, self.proj_embedding_size)
        self.geu = GatedEmbeddingUnit(self.proj_input, 1024, False)

    def pad(self, x):
        if x.shape[-1] == self.proj_input:
            return x
        return F.interpolate(x.reshape(-1, 1, x.shape[-1]), self.proj_input).reshape(*x.shape[:-1], self.proj_input)

    def forward(self, xs):
        if len(xs) < 2:
            raise ValueError("CollaborativeGating needs at least two experts")
        work = list(xs)
        total = None
        for _ in range(len(xs)):
            c = self.projection(self.pad(work.pop(0)))
            t = sum(c + self.projection(self.pad(o)) for o in work)
            gated = c * torch.sigmoid(c + self.projection(t))
            total = gated if total is None else total + gated
            work.append(c)
        B, S = total.shape[:2]
        return self.geu(total.reshape(B * S, -1)).reshape(B, S, -1)


# ------------------------------------------------------------------ FrameTransformer (two-stream model)
class FrameTransformer(nn.Module):
    """src/models/frame_transformer.py:83-244 with the widths as arguments and the CNN backbones injected.

    ``vid_model`` / ``img_model`` map the stacked raw clips / frames to one ``d``-wide feature vector each
    (VidResNet / ImgResNet, :50-74); ``None`` = FEATURE MODE: the inputs already are backbone features
    (``vid`` (B, S-1, d), ``img`` (B, S-1, d)) and the learned CLS inputs ``vid_cls`` / ``img_cls`` are (1, d) feature
    rows instead of a raw (1, 12, 3, 112, 112) clip / (1, 3, 224, 224) frame.  Everything after the backbone follows the
    reference statement by statement: per-clip CLS concat (:193-196, :213-216), ``view(batch_size, S, d)`` (:204, :222),
    seq-first PE + TransformerBase (:205-208, :227-231), CLS = token 0 (:209, :233), GELU head (:106).

    What the reference cannot run at this commit, and how it is read here (each restoration is the minimal one):
      * ``img_model`` / ``scene_transformer`` / ``img_cls`` are commented out (:94,98,104) but used by every mode except
        "vid": restored exactly as written in those comments (TransformerBase(d, d, 4, d, 4, 0.5)).
      * "sum": ``torch.cat((data, distil_inject))`` (:226) joins a (S, B, d) sequence with a (B, d) CLS vector — read as
        ``distil_inject.unsqueeze(0)``; the S + 1 tokens need ``max_len >= S + 1`` in the PositionalEncoding (:91-93).
      * "distil": ``forward`` returns the image stream's (CLS, last token) d-wide vectors and ``training_step`` feeds
        them to BCE against (B, n_classes) targets (:247-251) — a shape error; read as student = head(img CLS),
        teacher = head(video CLS) (the quantities the loss names ``img`` / ``vid`` suggest).
      * "pre_modal": ``vid = self.vid_step`` (:188, no call) and the mode string tested in ``img_step`` is "pre-modal"
        (:219), so no injection ever happens: what executes is identical to "frame".
      * "sum_residual" normalises ``img_cls`` twice and never uses ``vid_cls`` (:157-158): reproduced as written.
    """

    def __init__(self, model="vid", batch_size=2, seq_len=13, cls=1, d=896, n_classes=19, dropout=0.5, nlayers=4,
                 vid_model=None, img_model=None, **_unused):
        super().__init__()
        self.model, self.batch_size, self.d = model, batch_size, d
        self.seq_len = seq_len + (1 if cls else 0)                                   # :86-87
        self.criterion = nn.BCEWithLogitsLoss()
        self.distil_criterion = nn.CrossEntropyLoss()
        self.position_encoder = PositionalEncoding(d, dropout, max_len=self.seq_len + (1 if model == "sum" else 0))
        self.vid_model = vid_model
        self.distil_transformer = TransformerBase(d, 128, 2, 512, nlayers, dropout)  # :99
        self.vid_cls = nn.Parameter(torch.rand(1, d) if vid_model is None else torch.rand(1, 12, 3, 112, 112))
        self.img_mlp_head = nn.Sequential(nn.Linear(d, 512), nn.GELU(), nn.Linear(512, 128), nn.GELU(), nn.Linear(128, n_classes))
        self.norm = nn.LayerNorm(d)                                                  # :115 (unused)
        if model != "vid":
            self.img_model = img_model                                               # :94
            self.scene_transformer = TransformerBase(d, d, 4, d, nlayers, dropout)  # :98
            self.img_cls = nn.Parameter(torch.rand(1, d) if img_model is None else torch.rand(1, 3, 224, 224))  # :104

    def _stack(self, cls, data, backbone):
        total = [torch.cat((cls, data[i]), dim=0) for i in range(len(data))]         # :193-196
        x = torch.stack(total)
        if backbone is not None:
            x = backbone(x.view(-1, *x.shape[2:]))
        return x.reshape(self.batch_size, self.seq_len, self.d)

    def vid_step(self, data):
        x = self._stack(self.vid_cls, data, None if self.vid_model is None else
                        (lambda t: self.vid_model(t.permute(0, 2, 1, 3, 4))))       # :198-204
        x = self.distil_transformer(self.position_encoder(x.permute(1, 0, 2))).permute(1, 0, 2)
        return x[:, 0]

    def img_step(self, data, distil_inject):
        x = self._stack(self.img_cls, data, self.img_model).permute(1, 0, 2)          # :213-223
        if self.model == "sum":
            x = torch.cat((x, distil_inject.unsqueeze(0)))                            # :225-226
        seq = self.scene_transformer(self.position_encoder(x)).permute(1, 0, 2)
        cls = seq[:, 0]
        if self.model in ("distil", "sum"):
            return cls, seq[:, -1]                                                    # :234-239
        if self.model == "sum_residual":
            return cls, seq
        return self.img_mlp_head(cls)

    def forward(self, img, vid):
        m = self.model
        if m == "distil":
            vid_cls = self.vid_step(vid)
            img_cls, _ = self.img_step(img, vid_cls)
            return self.img_mlp_head(img_cls), self.img_mlp_head(vid_cls)
        if m == "sum":
            img_cls, vid_tkn = self.img_step(img, self.vid_step(vid))
            return self.img_mlp_head(img_cls + vid_tkn)                               # :143-147
        if m == "sum_residual":
            self.vid_step(vid)
            img_cls, _ = self.img_step(img, None)
            a = F.normalize(img_cls, p=2.0, dim=-1)
            b = F.normalize(a, p=2.0, dim=-1)                                         # :157-158 (img_cls again)
            return self.img_mlp_head(a + b)
        if m in ("frame", "pre_modal"):
            return self.img_step(img, None)                                           # :172-178
        if m == "vid":
            return self.img_mlp_head(self.vid_step(vid))                              # :179-182
        return None

    def loss(self, batch):
        """training_step, frame_transformer.py:246-282."""
        target, img, vid = batch
        if self.model == "distil":
            s, t = self(img, vid)
            return self.criterion(s, target.float()) + self.distil_criterion(s, torch.argmax(t, dim=-1))
        return self.criterion(self(img, vid), target.float())


class ViViT(nn.Module):
    """src/models/vit.py:79-128: patch embedding (einops Rearrange + Linear), per-frame space transformer over
    (space token + patches) with a learned (1, frames, patches + 1, dim) positional embedding, then a temporal transformer
    over (temporal token + per-frame CLS vectors), cls / mean pooling, LayerNorm + Linear head."""

    def __init__(self, image_size, patch_size, num_classes, num_frames, dim=192, depth=4, heads=3, pool="cls", in_channels=3,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, scale_dim=4):
        super().__init__()
        from einops.layers.torch import Rearrange
        num_patches = (image_size // patch_size) ** 2
        patch_dim = in_channels * patch_size ** 2
        self.to_patch_embedding = nn.Sequential(
            Rearrange("b t c (h p1) (w p2) -> b t (h w) (p1 p2 c)", p1=patch_size, p2=patch_size), nn.Linear(patch_dim, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames, num_patches + 1, dim))
        self.space_token = nn.Parameter(torch.randn(1, 1, dim))
        self.space_transformer = VitTransformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.temporal_token = nn.Parameter(torch.randn(1, 1, dim))
        self.temporal_transformer = VitTransformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.dropout = nn.Dropout(emb_dropout)
        self.pool = pool
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

    def forward(self, x):
        x = self.to_patch_embedding(x)
        b, t, n, d = x.shape
        x = torch.cat((self.space_token.expand(b, t, 1, d), x), dim=2)
        x = self.dropout(x + self.pos_embedding[:, :, :(n + 1)])
        x = self.space_transformer(x.reshape(b * t, n + 1, d))[:, 0].reshape(b, t, d)
        x = self.temporal_transformer(torch.cat((self.temporal_token.expand(b, 1, d), x), dim=1))
        x = x.mean(dim=1) if self.pool == "mean" else x[:, 0]
        return self.mlp_head(x)
