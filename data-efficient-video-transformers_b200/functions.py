"""torch.autograd.Functions over the C-ABI kernels.  Each Function is one block of the reference model
with a hand-scheduled backward (fused epilogues, recomputed dropout masks, no tensor saved twice):

  EmbedFn         SimpleTransformer.add_pos_cls               src/models/transformer.py:74-82
  EncoderLayerFn  nn.TransformerEncoderLayer (post-norm) and  torch/nn/modules/transformer.py:952-982
                  the cross-modal block (q from x, k/v from mem)
  LinearFn        nn.Linear (input projection)
  MlpFn           Linear -> act -> Dropout chains             src/models/TPN.py:88-96, frame_transformer.py:106
  LayerNormFn     nn.LayerNorm
  HeadLinearFn    Linear(d, n_classes)                        src/models/transformer.py:54
  ReadoutFn       CLS gather (+ expert sum) and sum_group     src/models/transformer.py:123-130, TPN.py:64-72
  DistillLossFn   BCE + CE(argmax teacher) + KL               src/models/frame_transformer.py:250-257
  PyramidHeadFn   sigmoid, mean over scales, BCE              src/models/TPN.py:98,112
  SpatialPoolFn   AvgPool2d to 1x1 of a CNN feature map       src/models/TPN.py:5,19,32
  CollabMixFn / CollabGateFn / L2NormFn   the glue of collaborative gating   src/models/collabgating.py:35-53,66-69,83-85

Tokens are kept batch-major inside the package: a (B, S, d) sequence batch is a [B*S, d] matrix.
"""
import math

import torch

from . import ops, ddp
from .capi import ACT_GELU, ACT_NONE, ACT_RELU

_seed_counter = [0]


def next_seed():
    """Fresh 63-bit dropout seed derived from torch's global seed and a call counter."""
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD6E8FEB86659FD93) & ((1 << 63) - 1)


def _zeros(n, dev):
    return torch.zeros(n, dtype=torch.float32, device=dev)


class _Sink:
    """Where a backward writes one parameter's gradient.  When the parameter's ``.grad`` is a view into a live gradient
    bucket (``ddp.GradBucketReducer(direct=True)``) the kernels accumulate straight into it and autograd receives
    ``None``; otherwise a zero-filled temporary is handed to autograd as usual."""

    def __init__(self, param, shape, dev):
        tgt = ddp.direct_target(param) if param is not None else None
        if tgt is not None and tuple(tgt[0].shape) == tuple(shape):
            self.buf, self._done, self.direct = tgt[0], tgt[1], True
        else:
            self.buf, self._done, self.direct = None, None, False
            self._shape, self._dev = shape, dev

    def zeros(self):
        """Accumulation target (zero-filled temporary, or the bucket view that zero_grad() cleared)."""
        if self.buf is None:
            self.buf = torch.zeros(self._shape, dtype=torch.float32, device=self._dev)
        return self.buf

    def result(self):
        """Call after the launches that write the gradient: what to return to autograd."""
        if self.direct:
            self._done()
            return None
        return self.buf


def _note_uses(ctx, params, first):
    """Forward-side registration of the parameters this Function's backward will write through ``_Sink``s (see
    ``ddp.note_use``); ``first`` = position of ``params[0]`` among forward's arguments."""
    for i, p in enumerate(params):
        if p is not None and ctx.needs_input_grad[first + i]:
            ddp.note_use(p)


class LayerCfg:
    """Static description of one attention block call."""

    def __init__(self, mode, batch, heads, dropout=0.0, training=False, activation="relu", attn_impl=0):
        self.mode, self.B, self.H = mode, batch, heads
        self.p = float(dropout) if training else 0.0
        self.act = {"relu": ACT_RELU, "gelu": ACT_GELU}[activation]
        self.attn_impl = attn_impl


class EncoderLayerFn(torch.autograd.Function):
    """x1 = LN1(x + Drop(OutProj(Attn(q(x), kv(mem or x)))));  out = LN2(x1 + Drop(W2 Drop(act(W1 x1))))."""

    @staticmethod
    def forward(ctx, cfg, x, mem, in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b):
        m = cfg.mode
        n, d = x.shape
        B, H = cfg.B, cfg.H
        hd, Sq, ff = d // H, n // B, l1_w.shape[0]
        p = cfg.p
        seeds = [next_seed() for _ in range(4)] if p > 0 else [0, 0, 0, 0]
        xp = m.fwd_planes(x)
        if mem is None:
            qkv = m.linear_fwd(xp, n, d, in_w, in_b)
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
            Sk, kv, memp = Sq, None, None
        else:
            nk = mem.shape[0]
            Sk = nk // B
            memp = m.fwd_planes(mem)
            qkv = m.linear_fwd(xp, n, d, in_w, in_b[:d], rows=(0, d))                 # q
            kv = m.linear_fwd(memp, nk, d, in_w, in_b[d:], rows=(d, 3 * d))          # [nk, 2d]
            q, k, v = qkv, kv[:, :d], kv[:, d:]
        scale = 1.0 / math.sqrt(hd)
        attn, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, hd, scale, dropout_p=p, seed=seeds[0], impl=cfg.attn_impl)
        attnp = m.fwd_planes(attn)
        y1 = m.linear_fwd(attnp, n, d, out_w, out_b, residual=x, dropout_p=p, seed=seeds[1])
        x1, mean1, rstd1 = ops.layernorm_fwd(y1, n1_w, n1_b)
        x1p = m.fwd_planes(x1)
        z = m.empty(n, ff, device=x.device) if cfg.act == ACT_GELU else None
        h = m.linear_fwd(x1p, n, d, l1_w, l1_b, act=cfg.act, dropout_p=p, seed=seeds[2], preact=z)
        hp = m.fwd_planes(h)
        y2 = m.linear_fwd(hp, n, ff, l2_w, l2_b, residual=x1, dropout_p=p, seed=seeds[3])
        out, mean2, rstd2 = ops.layernorm_fwd(y2, n2_w, n2_b)
        ctx.cfg, ctx.seeds, ctx.dims = cfg, seeds, (n, d, B, H, hd, Sq, Sk, ff)
        ctx.cross = mem is not None
        ctx.params = (in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b)   # gradient sinks (leaf parameters)
        _note_uses(ctx, ctx.params, 3)
        ctx.save_for_backward(x, mem, qkv, kv, attn, lse, y1, mean1, rstd1, x1, h, z, y2, mean2, rstd2,
                              in_w, out_w, l1_w, l2_w, n1_w, n2_w)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, mem, qkv, kv, attn, lse, y1, mean1, rstd1, x1, h, z, y2, mean2, rstd2,
         in_w, out_w, l1_w, l2_w, n1_w, n2_w) = ctx.saved_tensors
        cfg, seeds = ctx.cfg, ctx.seeds
        m, p = cfg.mode, cfg.p
        n, d, B, H, hd, Sq, Sk, ff = ctx.dims
        dev = x.device
        dout = dout.contiguous()
        P = ctx.params
        s_inw, s_inb = _Sink(P[0], (3 * d, d), dev), _Sink(P[1], (3 * d,), dev)
        s_ow, s_ob = _Sink(P[2], (d, d), dev), _Sink(P[3], (d,), dev)
        s_l1w, s_l1b = _Sink(P[4], (ff, d), dev), _Sink(P[5], (ff,), dev)
        s_l2w, s_l2b = _Sink(P[6], (d, ff), dev), _Sink(P[7], (d,), dev)
        s_n1w, s_n1b, s_n2w, s_n2b = (_Sink(P[i], (d,), dev) for i in (8, 9, 10, 11))

        def wgrad(sink, dyp, xp, M, N, K, bias_grad=None, dy=None):
            # bias_grad += colsum(dy): inside the weight-gradient GEMM when its kernel selection allows (ops.Mode.wgrad)
            if sink.direct:
                m.wgrad(dyp, xp, M, N, K, out=sink.buf, accumulate=True, bias_grad=bias_grad, dy=dy)
            else:
                sink.buf = m.wgrad(dyp, xp, M, N, K, bias_grad=bias_grad, dy=dy)

        # ---- LN2 and the feed-forward branch
        dy2, dz2 = ops.layernorm_bwd(dout, y2, mean2, rstd2, n2_w, dgamma=s_n2w.zeros(), dbeta=s_n2b.zeros(), dbias=s_l2b.zeros(),
                                     dropout_p=p, seed=seeds[3])
        dn2w, dn2b, dl2b = s_n2w.result(), s_n2b.result(), s_l2b.result()
        dz2p = m.split(dz2)
        wgrad(s_l2w, dz2p, m.split(h), n, d, ff)
        dl2w = s_l2w.result()
        dh = m.dgrad(dz2p, n, d, l2_w, relu_mask=h if cfg.act == ACT_RELU else None,
                     gelu_gate=z if cfg.act == ACT_GELU else None, dropout_p=p, seed=seeds[2])
        dhp = m.split(dh)
        wgrad(s_l1w, dhp, m.split(x1), n, ff, d, bias_grad=s_l1b.zeros(), dy=dh)
        dl1w, dl1b = s_l1w.result(), s_l1b.result()
        dx1 = m.dgrad(dhp, n, ff, l1_w, residual=dy2)
        # ---- LN1 and the attention branch
        dy1, dz1 = ops.layernorm_bwd(dx1, y1, mean1, rstd1, n1_w, dgamma=s_n1w.zeros(), dbeta=s_n1b.zeros(), dbias=s_ob.zeros(),
                                     dropout_p=p, seed=seeds[1])
        dn1w, dn1b, dob = s_n1w.result(), s_n1b.result(), s_ob.result()
        dz1p = m.split(dz1)
        wgrad(s_ow, dz1p, m.split(attn), n, d, d)
        dow = s_ow.result()
        dattn = m.dgrad(dz1p, n, d, out_w)
        scale = 1.0 / math.sqrt(hd)
        dinb_buf = s_inb.zeros()
        if not ctx.cross:
            dqkv = torch.empty_like(qkv)
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
            ops.attention_bwd(q, k, v, attn, dattn, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H, Sq, Sk,
                              hd, scale, dropout_p=p, seed=seeds[0], impl=cfg.attn_impl)
            dqkvp = m.split(dqkv)
            wgrad(s_inw, dqkvp, m.split(x), n, 3 * d, d, bias_grad=dinb_buf, dy=dqkv)
            dx = m.dgrad(dqkvp, n, 3 * d, in_w, residual=dy1) if ctx.needs_input_grad[1] else None
            dmem = None
        else:
            nk = mem.shape[0]
            dq = torch.empty_like(qkv)
            dkv = torch.empty_like(kv)
            ops.attention_bwd(qkv, kv[:, :d], kv[:, d:], attn, dattn, lse, dq, dkv[:, :d], dkv[:, d:], B, H, Sq, Sk,
                              hd, scale, dropout_p=p, seed=seeds[0], impl=cfg.attn_impl)
            dqp, dkvp = m.split(dq), m.split(dkv)
            if not s_inw.direct:
                s_inw.buf = torch.empty(3 * d, d, dtype=torch.float32, device=dev)
            m.wgrad(dqp, m.split(x), n, d, d, out=s_inw.buf[:d], accumulate=s_inw.direct, bias_grad=dinb_buf[:d], dy=dq)
            m.wgrad(dkvp, m.split(mem), nk, 2 * d, d, out=s_inw.buf[d:], accumulate=s_inw.direct, bias_grad=dinb_buf[d:], dy=dkv)
            dx = m.dgrad(dqp, n, d, in_w, residual=dy1, rows=(0, d)) if ctx.needs_input_grad[1] else None
            dmem = m.dgrad(dkvp, nk, 2 * d, in_w, rows=(d, 3 * d)) if ctx.needs_input_grad[2] else None
        dinw, dinb = s_inw.result(), s_inb.result()
        return (None, dx, dmem, dinw, dinb, dow, dob, dl1w, dl1b, dl2w, dl2b, dn1w, dn1b, dn2w, dn2b)


class PreNormLayerFn(torch.autograd.Function):
    """Pre-norm block of src/models/vit.py:8-75:  x1 = x + Drop(OutProj(Attn(LN1(x))));  out = x1 + Drop(W2 Drop(gelu(W1 LN2(x1)))).
    Bias-free packed qkv projection [3*inner, dim] (vit.py:39), scale = dim_head ** -0.5 (vit.py:37),
    inner = heads * dim_head may differ from dim; out_w is None when the reference uses nn.Identity (vit.py:41-44)."""

    @staticmethod
    def forward(ctx, cfg, dim_head, x, n1_w, n1_b, qkv_w, out_w, out_b, n2_w, n2_b, l1_w, l1_b, l2_w, l2_b):
        m = cfg.mode
        n, d = x.shape
        B, H = cfg.B, cfg.H
        S, inner, ff = n // B, H * dim_head, l1_w.shape[0]
        p = cfg.p
        seeds = [next_seed() for _ in range(3)] if p > 0 else [0, 0, 0]
        xn, mean1, rstd1 = ops.layernorm_fwd(x, n1_w, n1_b)
        qkv = m.linear_fwd(m.fwd_planes(xn), n, d, qkv_w, None)
        scale = dim_head ** -0.5
        attn, lse = ops.attention_fwd(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, H, S, S, dim_head, scale,
                                      impl=cfg.attn_impl)
        if out_w is not None:
            x1 = m.linear_fwd(m.fwd_planes(attn), n, inner, out_w, out_b, residual=x, dropout_p=p, seed=seeds[0])
        else:   # heads == 1 and dim_head == dim: no output projection, no dropout (vit.py:41-44)
            x1 = attn + x
        xn2, mean2, rstd2 = ops.layernorm_fwd(x1, n2_w, n2_b)
        z = m.empty(n, ff, device=x.device)
        h = m.linear_fwd(m.fwd_planes(xn2), n, d, l1_w, l1_b, act=ACT_GELU, dropout_p=p, seed=seeds[1], preact=z)
        out = m.linear_fwd(m.fwd_planes(h), n, ff, l2_w, l2_b, residual=x1, dropout_p=p, seed=seeds[2])
        ctx.cfg, ctx.seeds, ctx.dims = cfg, seeds, (n, d, B, H, dim_head, S, inner, ff)
        ctx.has_out = out_w is not None
        ctx.save_for_backward(x, xn, mean1, rstd1, qkv, attn, lse, x1, xn2, mean2, rstd2, z, h,
                              n1_w, qkv_w, out_w, n2_w, l1_w, l2_w)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, xn, mean1, rstd1, qkv, attn, lse, x1, xn2, mean2, rstd2, z, h,
         n1_w, qkv_w, out_w, n2_w, l1_w, l2_w) = ctx.saved_tensors
        cfg, seeds = ctx.cfg, ctx.seeds
        m, p = cfg.mode, cfg.p
        n, d, B, H, hd, S, inner, ff = ctx.dims
        dev = x.device
        dout = dout.contiguous()
        # ---- feed-forward branch: out = x1 + Drop(W2 h + b2)
        dz2 = ops.act_bwd(dout, dout, ACT_NONE, p, seeds[2]) if p > 0 else dout
        dl2b = _zeros(d, dev)
        dz2p = m.split(dz2)
        dl2w = m.wgrad(dz2p, m.split(h), n, d, ff, bias_grad=dl2b, dy=dz2)
        dzp = m.dgrad(dz2p, n, d, l2_w, gelu_gate=z, dropout_p=p, seed=seeds[1])      # d(pre-GELU)
        dl1b = _zeros(ff, dev)
        dzpp = m.split(dzp)
        dl1w = m.wgrad(dzpp, m.split(xn2), n, ff, d, bias_grad=dl1b, dy=dzp)
        dxn2 = m.dgrad(dzpp, n, ff, l1_w)
        dn2w, dn2b = _zeros(d, dev), _zeros(d, dev)
        dx1, _ = ops.layernorm_bwd(dxn2, x1, mean2, rstd2, n2_w, dgamma=dn2w, dbeta=dn2b, dres=dout)
        # ---- attention branch: x1 = x + Drop(Wo attn + bo)
        if ctx.has_out:
            dz1 = ops.act_bwd(dx1, dx1, ACT_NONE, p, seeds[0]) if p > 0 else dx1
            dob = _zeros(d, dev)
            dz1p = m.split(dz1)
            dow = m.wgrad(dz1p, m.split(attn), n, d, inner, bias_grad=dob, dy=dz1)
            dattn = m.dgrad(dz1p, n, d, out_w)
        else:
            dow, dob, dattn = None, None, dx1
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], attn, dattn, lse, dqkv[:, :inner],
                          dqkv[:, inner:2 * inner], dqkv[:, 2 * inner:], B, H, S, S, hd, hd ** -0.5, impl=cfg.attn_impl)
        dqkvp = m.split(dqkv)
        dqkvw = m.wgrad(dqkvp, m.split(xn), n, 3 * inner, d)
        dxn = m.dgrad(dqkvp, n, 3 * inner, qkv_w)
        dn1w, dn1b = _zeros(d, dev), _zeros(d, dev)
        dx, _ = ops.layernorm_bwd(dxn, x, mean1, rstd1, n1_w, dgamma=dn1w, dbeta=dn1b, dres=dx1)
        return (None, None, dx, dn1w, dn1b, dqkvw, dow, dob, dn2w, dn2b, dl1w, dl1b, dl2w, dl2b)


class EmbedFn(torch.autograd.Function):
    """tokens[b, s] = LN(Drop((s == 0 ? cls[b] : feat[b, s-1]) + pe[s])) -> [B*S, d]."""

    @staticmethod
    def forward(ctx, mode, feat, cls, pe, norm_w, norm_b, p):
        B, T, d = feat.shape
        # the kernel cannot know the buffers' lengths: refuse here what the reference refuses with a torch.cat / broadcast
        # error (a batch larger than hparams.batch_size CLS slots, a sequence longer than PositionalEncoding's max_len)
        if cls.numel() % d != 0 or B > cls.numel() // d:
            raise ValueError(f"embed: batch of {B} clips but the CLS parameter {tuple(cls.shape)} holds {cls.numel() // d} batch slots")
        if pe.numel() % d != 0 or T + 1 > pe.numel() // d:
            raise ValueError(f"embed: sequence of {T} + 1 tokens exceeds PositionalEncoding max_len {pe.numel() // d}")
        seed = next_seed() if p > 0 else 0
        cls_act = cls.detach().reshape(-1, d)[:B].to(mode.dtype).contiguous()
        pe2 = pe.reshape(-1, d)[: T + 1].contiguous()
        y, pre, mean, rstd = ops.embed_fwd(feat.contiguous(), cls_act, pe2, norm_w, norm_b, dropout_p=p, seed=seed)
        ctx.save_for_backward(pre, mean, rstd, norm_w)
        ctx.meta = (B, T + 1, d, p, seed, cls.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        pre, mean, rstd, norm_w = ctx.saved_tensors
        B, S, d, p, seed, cls_shape = ctx.meta
        dgw, dgb = _zeros(d, dy.device), _zeros(d, dy.device)
        dfeat, dcls = ops.embed_bwd(dy.contiguous(), pre, mean, rstd, norm_w, B, S, dgamma=dgw, dbeta=dgb,
                                    dropout_p=p, seed=seed, need_dfeat=ctx.needs_input_grad[1])
        dcls_full = torch.zeros(cls_shape, dtype=torch.float32, device=dy.device)
        dcls_full.reshape(-1, d)[:B] = dcls.float()
        return None, dfeat, dcls_full, None, dgw, dgb, None


class PosEncFn(torch.autograd.Function):
    """y = Drop(x + pe[s]) on batch-major tokens [B*S, d] (PositionalEncoding.forward, transformer.py:23-25)."""

    @staticmethod
    def forward(ctx, x, pe, S, p):
        seed = next_seed() if p > 0 else 0
        ctx.meta = (p, seed)
        return ops.posenc_fwd(x.contiguous(), pe, S, p, seed)

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.meta
        dx = ops.act_bwd(dy.contiguous(), dy, ACT_NONE, p, seed) if p > 0 else dy
        return dx, None, None, None


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tensor cores; x [M, K] (activation dtype of the mode)."""

    @staticmethod
    def forward(ctx, mode, x, w, b):
        M, K = x.shape
        xp = mode.fwd_planes(x)
        y = mode.linear_fwd(xp, M, K, w, b)
        ctx.mode = mode
        ctx.params = (w, b)
        _note_uses(ctx, ctx.params, 2)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        m = ctx.mode
        M, K = x.shape
        N = w.shape[0]
        dy = dy.contiguous()
        dyp = m.split(dy)
        s_w, s_b = _Sink(ctx.params[0], (N, K), dy.device), _Sink(ctx.params[1], (N,), dy.device)
        if s_w.direct:
            m.wgrad(dyp, m.split(x), M, N, K, out=s_w.buf, accumulate=True, bias_grad=s_b.zeros(), dy=dy)
        else:
            s_w.buf = m.wgrad(dyp, m.split(x), M, N, K, bias_grad=s_b.zeros(), dy=dy)
        dx = m.dgrad(dyp, M, N, w) if ctx.needs_input_grad[1] else None
        return None, dx, s_w.result(), s_b.result()


class MlpFn(torch.autograd.Function):
    """y_i = Drop_i(act_i(y_{i-1} W_i^T + b_i)) for a chain of tensor-core Linear layers (N % 8 == 0).
    acts / drops are per-layer lists.  Skinny-M, huge-K first layers run split-K (fp32 atomics) followed by
    a fused bias+act+dropout pass."""

    @staticmethod
    def forward(ctx, mode, acts, drops, x, *wb):
        nl = len(acts)
        ws, bs = wb[:nl], wb[nl:]
        M = x.shape[0]
        ys, zs, seeds = [], [], []
        cur = x
        for i in range(nl):
            K = cur.shape[1]
            N = ws[i].shape[0]
            p = drops[i]
            seed = next_seed() if p > 0 else 0
            seeds.append(seed)
            curp = mode.fwd_planes(cur)
            z = mode.empty(M, N, device=x.device) if acts[i] == ACT_GELU else None
            tiles = ((M + 127) // 128) * ((N + 127) // 128)
            kb = (K + 63) // 64
            if tiles * 4 <= ops.num_sms() and kb >= 16:
                exact = getattr(curp, "exact", False)
                wh, wl = mode.weight_exact(ws[i]) if exact else mode.weight(ws[i])
                Kx = 4 * K if exact else K
                acc = torch.zeros(M, N, dtype=torch.float32, device=x.device)
                splits = (ops.exact_splits(Kx) if exact else
                          max(2, ops.pick_splits(((M + 127) // 128) * ((N + 255) // 256), (Kx + 63) // 64, 64)))
                ops.gemm(curp[0], wh, M, N, Kx, a_lo=curp[1], b_lo=wl, out_f32=acc, splits=splits, atomic=True)
                if z is not None:
                    ops.bias_act(acc, bs[i], z, ACT_NONE)
                y = mode.empty(M, N, device=x.device)
                ops.bias_act(acc, bs[i], y, acts[i], p, seed)
            else:
                y = mode.linear_fwd(curp, M, K, ws[i], bs[i], act=acts[i], dropout_p=p, seed=seed, preact=z)
            ys.append(y)
            zs.append(z)
            cur = y
        ctx.mode, ctx.acts, ctx.drops, ctx.seeds, ctx.nl = mode, acts, drops, seeds, nl
        ctx.save_for_backward(x, *ys, *[z if z is not None else x.new_empty(0) for z in zs], *ws)
        ctx.params = (ws, bs)
        _note_uses(ctx, wb, 4)
        return cur

    @staticmethod
    def backward(ctx, dy):
        nl, m = ctx.nl, ctx.mode
        saved = ctx.saved_tensors
        x, ys, zs, ws = saved[0], saved[1:1 + nl], saved[1 + nl:1 + 2 * nl], saved[1 + 2 * nl:]
        acts, drops, seeds = ctx.acts, ctx.drops, ctx.seeds
        M = x.shape[0]
        dws, dbs = [None] * nl, [None] * nl
        # gradient wrt the last layer's pre-activation
        last = nl - 1
        if acts[last] != ACT_NONE or drops[last] > 0:
            ref = ys[last] if acts[last] != ACT_GELU else zs[last]
            dpre = ops.act_bwd(dy.contiguous(), ref, acts[last], drops[last], seeds[last])
        else:
            dpre = dy.contiguous()
        for i in range(last, -1, -1):
            inp = ys[i - 1] if i > 0 else x
            N, K = ws[i].shape
            dprep = m.split(dpre)
            s_w, s_b = _Sink(ctx.params[0][i], (N, K), dy.device), _Sink(ctx.params[1][i], (N,), dy.device)
            if s_w.direct:
                m.wgrad(dprep, m.split(inp), M, N, K, out=s_w.buf, accumulate=True, bias_grad=s_b.zeros(), dy=dpre)
            else:
                s_w.buf = m.wgrad(dprep, m.split(inp), M, N, K, bias_grad=s_b.zeros(), dy=dpre)
            dws[i], dbs[i] = s_w.result(), s_b.result()
            if i > 0:
                dpre = m.dgrad(dprep, M, N, ws[i], relu_mask=ys[i - 1] if acts[i - 1] == ACT_RELU else None,
                               gelu_gate=zs[i - 1] if acts[i - 1] == ACT_GELU else None,
                               dropout_p=drops[i - 1], seed=seeds[i - 1])
            elif ctx.needs_input_grad[3]:
                dpre = m.dgrad(dprep, M, N, ws[i])
            else:
                dpre = None
        return (None, None, None, dpre, *dws, *dbs)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x.contiguous(), w, b, eps)
        ctx.save_for_backward(x, mean, rstd, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, w = ctx.saved_tensors
        d = x.shape[1]
        dw, db = _zeros(d, dy.device), _zeros(d, dy.device)
        dx, _ = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, w, dgamma=dw, dbeta=db)
        return dx, dw, db, None


class HeadLinearFn(torch.autograd.Function):
    """logits[M, C] (fp32) = x W^T + b for a class head whose N is not tensor-core shaped."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return ops.head_linear_fwd(x.contiguous(), w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = ops.head_linear_bwd(x, w, dy.contiguous().float(), need_dx=ctx.needs_input_grad[0])
        return dx, dw, db


class ReadoutFn(torch.autograd.Function):
    """(tokens_0, ..., tokens_{E-1}) each [B*S, d]  ->  (sum_e CLS_e [B, d],  pooled_g ... for tokens_0).
    ``groups`` empty -> CLS readout only.  Backward builds every token gradient in one pass: pooled
    gradients land on the frame rows of tokens_0, the CLS gradient on row 0 of every expert."""

    @staticmethod
    def forward(ctx, B, S, groups, *tokens):
        d = tokens[0].shape[1]
        cls = ops.cls_sum([t.contiguous() for t in tokens], B, S, d)
        outs = ops.pyramid_pool_fwd(tokens[0], B, S, d, groups, relu=True) if groups else []
        ctx.meta = (B, S, d, tuple(groups), len(tokens))
        ctx.save_for_backward(*outs)
        return (cls, *outs)

    @staticmethod
    def backward(ctx, dcls, *douts):
        B, S, d, groups, E = ctx.meta
        outs = ctx.saved_tensors
        ref = dcls if dcls is not None else next(t for t in douts if t is not None)
        grads = []
        for e in range(E):
            if e == 0 and groups:
                g = torch.empty(B * S, d, dtype=ref.dtype, device=ref.device)
                dpool = [torch.zeros_like(o) if t is None else t.contiguous() for t, o in zip(douts, outs)]
                ops.pyramid_pool_bwd(dpool, outs, g, B, S, d, groups, relu=True)
            else:
                g = torch.zeros(B * S, d, dtype=ref.dtype, device=ref.device)
            if dcls is not None:
                g.view(B, S, d)[:, 0] = dcls
            else:
                g.view(B, S, d)[:, 0] = 0
            grads.append(g)
        return (None, None, None, *grads)


class DistillLossFn(torch.autograd.Function):
    """loss = w_bce BCEWithLogits(s, y) + w_ce CE(s, argmax t) + w_kl T^2 KL(softmax(t/T) || softmax(s/T)).
    Returns the 5-vector (total, bce, ce, kl, cos(s, t)[0]); only element 0 carries gradient."""

    @staticmethod
    def forward(ctx, student, teacher, target, w_bce, w_ce, w_kl, temperature):
        losses, dl = ops.distill_loss(student.contiguous(), None if teacher is None else teacher.contiguous(),
                                      target.contiguous().float(), w_bce=w_bce, w_ce=w_ce, w_kl=w_kl,
                                      temperature=temperature, need_grad=True)
        ctx.save_for_backward(dl)
        return losses

    @staticmethod
    def backward(ctx, dlosses):
        (dl,) = ctx.saved_tensors
        return dl * dlosses[0], None, None, None, None, None, None


class PyramidHeadFn(torch.autograd.Function):
    """z [G, B, C] fp32 scale logits -> (prob [B, C], bce loss [1]) with prob = mean_g sigmoid(z_g)."""

    @staticmethod
    def forward(ctx, z, target):
        prob, loss, dz = ops.pyramid_head(z.contiguous(), target.contiguous().float(), need_grad=True)
        ctx.save_for_backward(dz)
        ctx.mark_non_differentiable(prob)
        return prob, loss

    @staticmethod
    def backward(ctx, _dprob, dloss):
        (dz,) = ctx.saved_tensors
        return dz * dloss, None


class SpatialPoolFn(torch.autograd.Function):
    """maps [frames, C, H, W] (fp32 or bf16) -> pooled [frames, C] fp32 = mean over H*W (nn.AvgPool2d(kernel_size=H) on an
    H x H map, src/models/TPN.py:5,19,32); backward broadcasts dpooled / (H*W) over the map."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty(x.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
        ops.spatial_pool(x, out, 0)
        ctx.meta = (tuple(x.shape), x.dtype)
        return out

    @staticmethod
    def backward(ctx, dpooled):
        shape, dtype = ctx.meta
        dx = torch.empty(shape, dtype=dtype, device=dpooled.device)
        ops.spatial_pool_bwd(dpooled.contiguous().float(), dx, 0)
        return dx


class CollabMixFn(torch.autograd.Function):
    """T_i = (E-1) C_i + sum_{j>i} C_j + sum_{j<i} PC_j over stacked experts (collabgating.py:35-43,49)."""

    @staticmethod
    def forward(ctx, c, pc):
        return ops.collab_mix_fwd(c.contiguous(), pc.contiguous())

    @staticmethod
    def backward(ctx, dt):
        return ops.collab_mix_bwd(dt.contiguous())


class CollabGateFn(torch.autograd.Function):
    """g = sum_i C_i * sigmoid(C_i + A_i) (ContextGating's GLU, collabgating.py:83-85, summed over experts :50)."""

    @staticmethod
    def forward(ctx, c, a):
        c, a = c.contiguous(), a.contiguous()
        ctx.save_for_backward(c, a)
        return ops.collab_gate_fwd(c, a)

    @staticmethod
    def backward(ctx, dg):
        c, a = ctx.saved_tensors
        return ops.collab_gate_bwd(c, a, dg.contiguous())


class L2NormFn(torch.autograd.Function):
    """y = F.normalize(x) per row (GatedEmbeddingUnit, collabgating.py:66-69); y fp32."""

    @staticmethod
    def forward(ctx, x):
        y, inv = ops.l2norm_fwd(x.contiguous())
        ctx.save_for_backward(y, inv)
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(dy.contiguous().float(), y, inv, ctx.in_dtype)
