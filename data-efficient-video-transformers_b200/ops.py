"""Functional layer over the C-ABI: every function takes/returns torch CUDA tensors and launches
hand-written sm_100a kernels from libtvt_b200.so on torch's current stream.  No function here has a
PyTorch/CPU fallback: a missing library or a non-CUDA tensor raises.

Precision modes (``Mode``):
  "bf16": activations and GEMM operands bf16, fp32 accumulation / statistics / losses, fp32 master weights.
  "fp32": activations fp32; every GEMM operand is split into two bf16 planes (hi + lo) and the tensor
          cores run hi*hi + hi*lo + lo*hi — the "fp32-accumulate" parity mode (1e-3 tolerance).
"""
import os
import weakref

import torch

from . import capi
from .capi import ACT_GELU, ACT_NONE, ACT_RELU, TVT_BF16, TVT_F32, TvtError

_NUM_SMS = None


def num_sms():
    global _NUM_SMS
    if _NUM_SMS is None:
        _NUM_SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _NUM_SMS


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _dt(t):
    if t.dtype == torch.float32:
        return TVT_F32
    if t.dtype == torch.bfloat16:
        return TVT_BF16
    raise TvtError(f"unsupported dtype {t.dtype}: the kernels take bf16 or fp32")


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise TvtError("tensor is not on a CUDA device: this path has no CPU implementation")


def _rowmajor(t):
    if t.stride(-1) != 1:
        raise TvtError("innermost dimension must be contiguous")
    return t.stride(0)


# ------------------------------------------------------------------------------------------ GEMM
def gemm(a, b, m, n, k, *, a_lo=None, b_lo=None, a_mn=False, b_mn=False, lda=None, ldb=None, bias=None,
         act=ACT_NONE, alpha=1.0, residual=None, relu_mask=None, gelu_gate=None, dropout_p=0.0, seed=0,
         out_f32=None, out_bf16=None, out_bf16_lo=None, out_preact=None, splits=1, atomic=False,
         ln_in=None, ln_res=None, stats_out=None, ln_dim=0, ln_eps=1e-5, a_rowsum=None):
    """``ln_in=(stats, c)`` / ``ln_res=(stats, gamma, beta)`` / ``stats_out``: the LayerNorm-folded inference epilogues;
    ``a_rowsum`` [m] fp32 += sum over k of A (the bias gradient inside a wgrad GEMM) - see tvt_gemm_args in include/tvt.h."""
    _cuda(a, b, a_lo, b_lo, bias, residual, relu_mask, gelu_gate, out_f32, out_bf16, out_bf16_lo, out_preact)
    g = capi.GemmArgs()
    g.a, g.a_lo, g.b, g.b_lo = _p(a), _p(a_lo), _p(b), _p(b_lo)
    g.m, g.n, g.k = m, n, k
    g.lda = lda if lda is not None else _rowmajor(a)
    g.ldb = ldb if ldb is not None else _rowmajor(b)
    g.a_mn_major, g.b_mn_major, g.splits, g.act, g.alpha = int(a_mn), int(b_mn), splits, act, alpha
    g.bias = _p(bias)
    if residual is not None:
        g.residual, g.residual_dtype, g.ld_residual = _p(residual), _dt(residual), _rowmajor(residual)
    if relu_mask is not None:
        g.relu_mask, g.mask_dtype, g.ld_mask = _p(relu_mask), _dt(relu_mask), _rowmajor(relu_mask)
    if gelu_gate is not None:
        g.gelu_gate, g.gate_dtype, g.ld_gate = _p(gelu_gate), _dt(gelu_gate), _rowmajor(gelu_gate)
    g.dropout_p, g.dropout_seed = dropout_p, seed
    if out_preact is not None:
        g.out_preact, g.preact_dtype, g.ld_preact = _p(out_preact), _dt(out_preact), _rowmajor(out_preact)
    if out_f32 is not None:
        g.out_f32, g.ld_f32 = _p(out_f32), _rowmajor(out_f32)
    g.atomic_out = int(atomic)
    if out_bf16 is not None:
        g.out_bf16, g.out_bf16_lo, g.ld_bf16 = _p(out_bf16), _p(out_bf16_lo), _rowmajor(out_bf16)
    if ln_in is not None:
        _cuda(*ln_in)
        g.ln_in_stats, g.ln_in_c = _p(ln_in[0]), _p(ln_in[1])
    if ln_res is not None:
        _cuda(*ln_res)
        g.ln_res_stats, g.ln_res_gamma, g.ln_res_beta = _p(ln_res[0]), _p(ln_res[1]), _p(ln_res[2])
    if stats_out is not None:
        _cuda(stats_out)
        g.stats_out = _p(stats_out)
    g.ln_dim, g.ln_eps = ln_dim, ln_eps
    if a_rowsum is not None:
        _cuda(a_rowsum)
        g.a_rowsum = _p(a_rowsum)
    capi.call("tvt_gemm", g, _stream())


def ln_fold_supported(m, n, k):
    return bool(capi.load().tvt_gemm_ln_fold_supported(m, n, k))


FUSE_BIAS_GRAD = os.environ.get("TVT_FUSE_BIAS_GRAD", "1") != "0"   # bias gradients inside the wgrad GEMMs (tvt_gemm_args.a_rowsum)
_rowsum_ok = {}


def rowsum_supported(m, n, k, splits):
    key = (m, n, k, splits)
    r = _rowsum_ok.get(key)
    if r is None:
        r = _rowsum_ok[key] = bool(capi.load().tvt_gemm_rowsum_supported(m, n, k, splits))
    return r


def split_f32(x, hi, lo=None):
    _cuda(x, hi, lo)
    s = capi.SplitArgs(_p(x), _p(hi), _p(lo), x.numel())
    capi.call("tvt_split_f32", s, _stream())


def split_f32x3(x, operand):
    """x fp32 [rows, cols] (row pitch x.stride(0)) -> (hi4, lo4) bf16 [rows, 4 * cols]: the K-concatenated three-plane
    layout of tvt_split_f32x3 (operand 0 = activation / A side, 1 = weight / B side)."""
    _cuda(x)
    rows, cols = x.shape
    hi4 = torch.empty(rows, 4 * cols, dtype=torch.bfloat16, device=x.device)
    lo4 = torch.empty_like(hi4)
    a = capi.Split3Args(_p(x), _p(hi4), _p(lo4), rows, cols, _rowmajor(x), operand, 0)
    capi.call("tvt_split_f32x3", a, _stream())
    return hi4, lo4


def exact_splits(k_total):
    """Split-K factor that keeps every tensor-core accumulation chain of an exact fp32-mode GEMM at <= 2 k-blocks of 64."""
    kb = (k_total + 63) // 64
    return max(1, (kb + 1) // 2)


class ExactPlanes(tuple):
    """(hi4, lo4) forward operand of the fp32 mode's exact GEMMs (see Mode.fwd_planes)."""
    exact = True


def colsum(x, out):
    """out[c] += sum_r x[r, c]; out fp32, zero-filled by the caller."""
    _cuda(x, out)
    a = capi.ColsumArgs(_p(x), _p(out), x.shape[0], x.shape[1], _rowmajor(x), _dt(x))
    capi.call("tvt_colsum", a, _stream())


def bias_act(x, bias, out, act=ACT_NONE, dropout_p=0.0, seed=0, residual=None, preact=None):
    """out = dropout(act(x + bias)) (+ residual); ``preact`` receives x + bias.  x fp32 [rows, cols] contiguous; residual /
    preact contiguous tensors of out's dtype."""
    _cuda(x, bias, out, residual, preact)
    for t in (residual, preact):
        if t is not None and (t.dtype != out.dtype or not t.is_contiguous()):
            raise TvtError("bias_act: residual / preact must be contiguous tensors of the output dtype")
    a = capi.BiasActArgs(_p(x), _p(bias), _p(out), x.shape[0], x.shape[1], _dt(out), act, dropout_p, seed, _p(residual), _p(preact))
    capi.call("tvt_bias_act_fwd", a, _stream())


def posenc_fwd(x, pe, S, dropout_p=0.0, seed=0):
    """x [B*S, d] batch-major tokens, pe fp32 [S, d] -> dropout(x + pe[s])."""
    _cuda(x, pe)
    y = torch.empty_like(x)
    a = capi.PosencArgs(_p(x), _p(pe), _p(y), x.shape[0], x.shape[1], S, _dt(x), dropout_p, seed)
    capi.call("tvt_posenc_fwd", a, _stream())
    return y


def act_bwd(dy, y_or_z, act, dropout_p=0.0, seed=0):
    _cuda(dy, y_or_z)
    dx = torch.empty_like(dy)
    a = capi.ActBwdArgs(_p(dy), _p(y_or_z), _p(dx), dy.shape[0], dy.shape[1], _dt(dy), act, dropout_p, seed)
    capi.call("tvt_act_bwd", a, _stream())
    return dx


def pick_splits(tiles, k_blocks, max_splits=32):
    """Split-K factor for a GEMM with few output tiles and a long contraction (wgrad, skinny MLP layers):
    minimise  waves x (k-blocks per split x 512 cycles + 4000 cycles of atomic epilogue)."""
    nsm = num_sms()
    best, best_cost = 1, None
    for s in range(1, max(1, min(max_splits, k_blocks)) + 1):
        waves = -(-tiles * s // nsm)
        cost = waves * (-(-k_blocks // s) * 512 + (4000 if s > 1 else 3000))
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


class Mode:
    """Precision mode + per-step cache of bf16 weight planes (keyed on the parameter's version counter)."""

    def __init__(self, name="bf16"):
        if name not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {name!r}")
        self.name = name
        self.fp32 = name == "fp32"
        self.dtype = torch.float32 if self.fp32 else torch.bfloat16
        self.tvt = TVT_F32 if self.fp32 else TVT_BF16
        self._wcache = {}

    def split(self, x):
        """Activation [rows, cols] -> (hi, lo) bf16 GEMM operand planes."""
        if not self.fp32:
            if x.dtype != torch.bfloat16:
                raise TvtError("bf16 mode expects bf16 activations")
            return x, None
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty_like(hi)
        split_f32(x.contiguous(), hi, lo)
        return hi, lo

    def fwd_planes(self, x):
        """Operand planes of a FORWARD GEMM.  bf16 mode: the activation itself.  fp32 mode: the three-plane K-concatenated
        layout, so that pre-activations carry fp32 accuracy (2^-24, not the 2^-17 of two planes) and no ReLU / GELU gate
        lands on the wrong side of zero relative to an fp32 reference; the backward GEMMs (dgrad / wgrad), whose errors
        are not amplified by a gate, keep the two-plane 3-pass form."""
        if not self.fp32:
            return self.split(x)
        if x.dim() != 2 or x.shape[1] % 8 != 0 or x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0:
            return self.split(x)
        return ExactPlanes(split_f32x3(x, 0))

    def weight_exact(self, w, rows=None):
        """fp32 master weight [N, K] -> cached (hi4, lo4) [N, 4K] planes of the exact forward GEMMs (see fwd_planes)."""
        key = (id(w), "x3")
        ent = self._wcache.get(key)
        if ent is None or ent[0]() is not w or ent[1] != w._version:
            wd = w.detach()
            if wd.dtype != torch.float32:
                raise TvtError("master weights must be fp32")
            hi4, lo4 = split_f32x3(wd.contiguous(), 1)
            ent = (weakref.ref(w), w._version, hi4, lo4)
            self._wcache[key] = ent
        hi, lo = ent[2], ent[3]
        if rows is not None:
            hi, lo = hi[rows[0]:rows[1]], lo[rows[0]:rows[1]]
        return hi, lo

    def folded_weight(self, w, b, gamma, beta):
        """A Linear fed by LayerNorm(gamma, beta), folded for the LN-free inference path (tvt_gemm_args.ln_in_*):
        (W' = bf16(W diag(gamma)) [N, K], c = rowsum(W') fp32 [N], b' = b + W beta fp32 [N]); cached per weight version.
        One-off host-side preparation (the teacher's weights are frozen), not part of the step."""
        key = (id(w), "lnfold")
        vers = (w._version, b._version, gamma._version, beta._version, id(gamma))
        ent = self._wcache.get(key)
        if ent is None or ent[0]() is not w or ent[1] != vers:
            with torch.no_grad():
                wf, g, bt = w.detach().float(), gamma.detach().float(), beta.detach().float()
                wp = (wf * g[None, :]).to(torch.bfloat16).contiguous()
                c = wp.float().sum(dim=1).contiguous()
                bp = (b.detach().float() + wf @ bt).contiguous()
            ent = (weakref.ref(w), vers, wp, c, bp, (weakref.ref(b), weakref.ref(gamma), weakref.ref(beta)))
            self._wcache[key] = ent
        return ent[2], ent[3], ent[4]

    def any_stale(self):
        """True when a cached operand (bf16 planes, three-plane layout, LayerNorm-folded weight) no longer matches the version
        of the tensors it was derived from - i.e. the next call would rebuild it at a NEW address.  CUDA graphs that captured the
        old addresses must be re-captured (hostapi.GraphedForward)."""
        for key, ent in self._wcache.items():
            w = ent[0]()
            if w is None:
                continue
            if isinstance(key, tuple) and key[1] == "lnfold":
                srcs = [r() for r in ent[5]]
                if any(t is None for t in srcs) or ent[1] != (w._version, srcs[0]._version, srcs[1]._version, srcs[2]._version, id(srcs[1])):
                    return True
            elif ent[1] != w._version:
                return True
        return False

    def weight(self, w, rows=None):
        """fp32 master weight [N, K] -> cached (hi, lo) bf16 planes, refreshed when the parameter changes
        (version counter) — one conversion per optimizer step.  ``rows=(r0, r1)`` selects a row slice
        (e.g. the q / kv parts of a packed in_proj_weight)."""
        ent = self._wcache.get(id(w))
        if ent is None or ent[0]() is not w or ent[1] != w._version:
            wd = w.detach()
            if wd.dtype != torch.float32:
                raise TvtError("master weights must be fp32")
            hi = torch.empty(wd.shape, dtype=torch.bfloat16, device=wd.device)
            lo = torch.empty_like(hi) if self.fp32 else None
            split_f32(wd.contiguous(), hi, lo)
            ent = (weakref.ref(w), w._version, hi, lo)
            self._wcache[id(w)] = ent
        hi, lo = ent[2], ent[3]
        if rows is not None:
            hi = hi[rows[0]:rows[1]]
            lo = lo[rows[0]:rows[1]] if lo is not None else None
        return hi, lo

    def empty(self, *shape, device):
        return torch.empty(*shape, dtype=self.dtype, device=device)

    # y[M,N] = act(x W^T + b) (+dropout) (+residual)
    def linear_fwd(self, xp, M, K, W, b, *, act=ACT_NONE, residual=None, dropout_p=0.0, seed=0, preact=None,
                   rows=None):
        if getattr(xp, "exact", False):
            # fp32 parity mode: three-plane operands (all 24 mantissa bits) AND short tensor-core accumulation chains.  The
            # fp32 accumulator in TMEM truncates when it aligns each MMA's products, a small BIASED error per instruction
            # that grows linearly with the chain (measured, tools/diag_gemm_precision.py: 1.7e-5 over K = 2048 in one chain
            # vs 2.8e-7 in chains of two 64-wide k-blocks, cuBLAS fp32: 2.9e-7), so the contraction is split into chains of
            # at most two k-blocks whose partial sums meet in fp32 atomics, and the epilogue runs as its own kernel.
            if W.shape[1] % 8 != 0:
                raise TvtError("exact forward planes need K % 8 == 0")
            wh, wl = self.weight_exact(W, rows)
            N = wh.shape[0]
            Kx = 4 * K
            acc = torch.zeros(M, N, dtype=torch.float32, device=xp[0].device)
            gemm(xp[0], wh, M, N, Kx, a_lo=xp[1], b_lo=wl, lda=_rowmajor(xp[0]), ldb=Kx, out_f32=acc,
                 splits=exact_splits(Kx), atomic=True)
            y = self.empty(M, N, device=xp[0].device)
            if residual is not None and not residual.is_contiguous():
                residual = residual.contiguous()
            bias_act(acc, b, y, act, dropout_p, seed, residual=residual, preact=preact)
            return y
        wh, wl = self.weight(W, rows)
        N = wh.shape[0]
        y = self.empty(M, N, device=xp[0].device)
        gemm(xp[0], wh, M, N, K, a_lo=xp[1], b_lo=wl, lda=_rowmajor(xp[0]), ldb=K, bias=b, act=act,
             residual=residual, dropout_p=dropout_p, seed=seed, out_preact=preact,
             out_f32=y if self.fp32 else None, out_bf16=None if self.fp32 else y)
        return y

    # dx[M,K] = dy[M,N] W[N,K]  (* relu mask, * gelu', dropout, + residual)
    def dgrad(self, dyp, M, N, W, *, residual=None, relu_mask=None, gelu_gate=None, dropout_p=0.0, seed=0,
              rows=None):
        K = W.shape[1]
        wh, wl = self.weight(W, rows)
        dx = self.empty(M, K, device=dyp[0].device)
        alpha = 1.0
        if relu_mask is not None and dropout_p > 0.0:
            # relu_mask is the activation AFTER forward dropout: dropped elements are already zero in it, so the mask test
            # covers both gates and the dropout stage reduces to its constant 1 / (1 - p) (no hash in the epilogue)
            thr16 = int(dropout_p * 65536.0 + 0.5)
            alpha, dropout_p = 65536.0 / (65536.0 - thr16), 0.0
        gemm(dyp[0], wh, M, K, N, a_lo=dyp[1], b_lo=wl, lda=_rowmajor(dyp[0]), ldb=_rowmajor(wh), b_mn=True,
             residual=residual, relu_mask=relu_mask, gelu_gate=gelu_gate, dropout_p=dropout_p, seed=seed, alpha=alpha,
             out_f32=dx if self.fp32 else None, out_bf16=None if self.fp32 else dx)
        return dx

    # dW[N,K] = dy[M,N]^T x[M,K]  (fp32, split-K over the token dimension when the tile grid is small)
    def wgrad(self, dyp, xp, M, N, K, out=None, accumulate=False, bias_grad=None, dy=None):
        """dW[N, K] = dY^T X.  ``accumulate=True`` adds into ``out`` (fp32 atomics) instead of overwriting it.
        ``bias_grad`` [N] fp32 (zeroed or accumulating) += colsum(dY): inside this GEMM when the kernel selection allows it
        (bf16 mode, split-K / accumulating launch on the CTA-pair tiles), else by the column-sum kernel over ``dy`` (the
        unsplit gradient tensor)."""
        splits = pick_splits(((N + 127) // 128) * ((K + 255) // 256), (M + 63) // 64)
        if out is None:
            out = (torch.zeros if splits > 1 else torch.empty)(N, K, dtype=torch.float32, device=dyp[0].device)
        elif splits > 1 and not accumulate:
            out.zero_()
        atomic = splits > 1 or accumulate
        # measured (tools/wgrad_time.py): the N = 16 row-sum MMAs cost ~19 % of the GEMM, the column-sum kernel one read of dY:
        # fused wins while in_features <= ~1200 (and always saves a launch)
        fused = (bias_grad is not None and FUSE_BIAS_GRAD and atomic and dyp[1] is None and K <= 1024
                 and rowsum_supported(N, K, M, splits))
        if bias_grad is not None and not fused:
            if dy is None and dyp[1] is not None:
                raise TvtError("wgrad: bias_grad in the fp32 mode needs the unsplit gradient tensor (dy=...)")
            colsum(dy if dy is not None else dyp[0], bias_grad)
        gemm(dyp[0], xp[0], N, K, M, a_lo=dyp[1], b_lo=xp[1], a_mn=True, b_mn=True, lda=_rowmajor(dyp[0]),
             ldb=_rowmajor(xp[0]), out_f32=out, splits=splits, atomic=atomic, a_rowsum=bias_grad if fused else None)
        return out


# ------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x, gamma, beta, eps=1e-5, *, save_stats=True, out=None):
    """``out=(buf [B, n_blocks * S, d], block)``: write the rows of clip b as rows [block * S, (block + 1) * S) of buf[b] (the
    blocked output of tvt_layernorm_fwd) and return that strided view [B, S, d] instead of a dense [rows, d] tensor."""
    _cuda(x, gamma, beta)
    rows, d = x.shape
    y_seq = y_pitch = 0
    if out is not None:
        buf, block = out
        _cuda(buf)
        B, tot, d2 = buf.shape
        S = rows // B
        if d2 != d or rows != B * S or tot % S or not 0 <= block < tot // S or not buf.is_contiguous() or buf.dtype != x.dtype:
            raise TvtError("layernorm_fwd: out buffer does not match [B, n_blocks * S, d]")
        y = buf[:, block * S:(block + 1) * S]
        y_seq, y_pitch = S, tot * d
    else:
        y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    a = capi.LayerNormFwdArgs()
    a.x, a.gamma, a.beta, a.y, a.mean, a.rstd = _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd)
    a.rows, a.d, a.seq_len, a.dtype, a.eps = rows, d, 0, _dt(x), eps
    a.y_seq, a.y_pitch = y_seq, y_pitch
    capi.call("tvt_layernorm_fwd", a, _stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dgamma=None, dbeta=None, dbias=None, dropout_p=0.0, seed=0, dres=None):
    """Returns (dx, dz): dz is the dropout-masked branch gradient (dz is dx when dropout_p == 0).
    ``dres`` (pre-norm blocks) is added to dx."""
    _cuda(dy, x, mean, rstd, gamma, dgamma, dbeta, dbias, dres)
    rows, d = x.shape
    dx = torch.empty_like(x)
    dz = torch.empty_like(x) if dropout_p > 0 else None
    a = capi.LayerNormBwdArgs()
    a.dy, a.x, a.mean, a.rstd, a.gamma, a.dx, a.dz = _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dx), _p(dz)
    a.dgamma, a.dbeta, a.dbias = _p(dgamma), _p(dbeta), _p(dbias)
    a.rows, a.d, a.seq_len, a.dtype, a.dropout_p, a.dropout_seed = rows, d, 0, _dt(x), dropout_p, seed
    a.dres = _p(dres)
    capi.call("tvt_layernorm_bwd", a, _stream())
    return dx, (dz if dz is not None else dx)


def embed_fwd(feat, cls, pe, gamma, beta, eps=1e-5, dropout_p=0.0, seed=0):
    """feat [B, T, d], cls [B, d], pe [S, d] fp32 -> tokens [B*S, d], pre-LN rows, mean, rstd."""
    _cuda(feat, cls, pe, gamma, beta)
    B, T, d = feat.shape
    S = T + 1
    y = torch.empty(B * S, d, dtype=feat.dtype, device=feat.device)
    pre = torch.empty_like(y)
    mean = torch.empty(B * S, dtype=torch.float32, device=feat.device)
    rstd = torch.empty_like(mean)
    a = capi.LayerNormFwdArgs()
    a.x, a.cls, a.pe, a.gamma, a.beta, a.y, a.pre, a.mean, a.rstd = (_p(feat), _p(cls), _p(pe), _p(gamma), _p(beta),
                                                                      _p(y), _p(pre), _p(mean), _p(rstd))
    a.rows, a.d, a.seq_len, a.dtype, a.eps, a.dropout_p, a.dropout_seed = B * S, d, S, _dt(feat), eps, dropout_p, seed
    capi.call("tvt_layernorm_fwd", a, _stream())
    return y, pre, mean, rstd


def embed_bwd(dy, pre, mean, rstd, gamma, B, S, *, dgamma, dbeta, dropout_p=0.0, seed=0, need_dfeat=True):
    d = pre.shape[1]
    dfeat = torch.empty(B, S - 1, d, dtype=pre.dtype, device=pre.device) if need_dfeat else None
    dcls = torch.empty(B, d, dtype=pre.dtype, device=pre.device)
    a = capi.LayerNormBwdArgs()
    a.dy, a.x, a.mean, a.rstd, a.gamma = _p(dy), _p(pre), _p(mean), _p(rstd), _p(gamma)
    a.dfeat, a.dcls, a.dgamma, a.dbeta = _p(dfeat), _p(dcls), _p(dgamma), _p(dbeta)
    a.rows, a.d, a.seq_len, a.dtype, a.dropout_p, a.dropout_seed = B * S, d, S, _dt(pre), dropout_p, seed
    capi.call("tvt_layernorm_bwd", a, _stream())
    return dfeat, dcls


# ------------------------------------------------------------------------------------------ attention
def attention_fwd(q, k, v, B, H, Sq, Sk, hd, scale, *, dropout_p=0.0, seed=0, impl=0):
    """q [B*Sq, >=H*hd] (row pitch q.stride(0)), k/v [B*Sk, ...]; returns o [B*Sq, H*hd], lse [B,H,Sq]."""
    _cuda(q, k, v)
    o = torch.empty(B * Sq, H * hd, dtype=q.dtype, device=q.device)
    lse = torch.empty(B, H, Sq, dtype=torch.float32, device=q.device)
    a = capi.AttentionFwdArgs()
    a.q, a.k, a.v, a.o, a.lse = _p(q), _p(k), _p(v), _p(o), _p(lse)
    a.batch, a.heads, a.sq, a.sk, a.head_dim = B, H, Sq, Sk, hd
    a.ldq, a.ldk, a.ldv, a.ldo = _rowmajor(q), _rowmajor(k), _rowmajor(v), H * hd
    a.scale, a.dtype, a.impl, a.dropout_p, a.dropout_seed = scale, _dt(q), impl, dropout_p, seed
    capi.call("tvt_attention_fwd", a, _stream())
    return o, lse


def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Sq, Sk, hd, scale, *, dropout_p=0.0, seed=0, impl=0):
    _cuda(q, k, v, o, do, lse, dq, dk, dv)
    a = capi.AttentionBwdArgs()
    a.q, a.k, a.v, a.o, a.d_o, a.lse, a.dq, a.dk, a.dv = (_p(q), _p(k), _p(v), _p(o), _p(do), _p(lse), _p(dq),
                                                          _p(dk), _p(dv))
    a.batch, a.heads, a.sq, a.sk, a.head_dim = B, H, Sq, Sk, hd
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = _rowmajor(q), _rowmajor(k), _rowmajor(v), _rowmajor(o), _rowmajor(do)
    a.lddq, a.lddk, a.lddv = _rowmajor(dq), _rowmajor(dk), _rowmajor(dv)
    a.scale, a.dtype, a.impl, a.dropout_p, a.dropout_seed = scale, _dt(q), impl, dropout_p, seed
    capi.call("tvt_attention_bwd", a, _stream())


# ------------------------------------------------------------------------------------------ pooling
def pyramid_pool_fwd(tokens, B, S, d, groups, relu=True, skip_cls=True):
    """tokens [B*S, d] -> list of out_g [B, floor(T/g)*d] over the frame tokens (CLS row skipped)."""
    _cuda(tokens)
    T = S - 1 if skip_cls else S
    a = capi.PyramidPoolFwdArgs()
    base = tokens.data_ptr() + (d * tokens.element_size() if skip_cls else 0)
    a.x, a.batch, a.frames, a.d, a.x_batch_stride, a.x_frame_stride = base, B, T, d, S * d, d
    a.num_scales, a.dtype, a.relu = len(groups), _dt(tokens), int(relu)
    outs = []
    for i, g in enumerate(groups):
        o = torch.empty(B, (T // g) * d, dtype=tokens.dtype, device=tokens.device)
        outs.append(o)
        a.groups[i] = g
        a.out[i] = o.data_ptr()
    capi.call("tvt_pyramid_pool_fwd", a, _stream())
    return outs


def pyramid_pool_bwd(douts, outs, dtokens, B, S, d, groups, relu=True, skip_cls=True, accumulate=False):
    """Writes (or accumulates) the frame-token rows of dtokens [B*S, d]."""
    _cuda(dtokens, *douts)
    T = S - 1 if skip_cls else S
    a = capi.PyramidPoolBwdArgs()
    base = dtokens.data_ptr() + (d * dtokens.element_size() if skip_cls else 0)
    a.dx, a.batch, a.frames, a.d, a.dx_batch_stride, a.dx_frame_stride = base, B, T, d, S * d, d
    a.num_scales, a.dtype, a.relu, a.accumulate = len(groups), _dt(dtokens), int(relu), int(accumulate)
    for i, g in enumerate(groups):
        a.groups[i] = g
        a.dout[i] = douts[i].data_ptr()
        a.out[i] = outs[i].data_ptr()
    capi.call("tvt_pyramid_pool_bwd", a, _stream())


def spatial_pool(x, out, col_offset):
    """x [frames, C, H, W] -> out[:, col_offset:col_offset+C] = mean over H*W."""
    _cuda(x, out)
    frames, Cc = x.shape[0], x.shape[1]
    a = capi.SpatialPoolArgs(_p(x), _p(out), frames, Cc, x.shape[2] * x.shape[3], _rowmajor(out), col_offset, _dt(x), _dt(out))
    capi.call("tvt_spatial_pool_fwd", a, _stream())


def spatial_pool_bwd(dpooled, dx, col_offset):
    """dx [frames, C, H, W] = dpooled[:, col_offset:col_offset+C] / (H*W) broadcast over the map; dpooled fp32."""
    _cuda(dpooled, dx)
    if dpooled.dtype != torch.float32:
        raise TvtError("spatial_pool_bwd: dpooled must be fp32")
    frames, Cc = dx.shape[0], dx.shape[1]
    a = capi.SpatialPoolBwdArgs(_p(dpooled), _p(dx), frames, Cc, dx.shape[2] * dx.shape[3], _rowmajor(dpooled), col_offset, _dt(dx), 0)
    capi.call("tvt_spatial_pool_bwd", a, _stream())


# ------------------------------------------------------------------------------------------ heads / loss
def head_linear_fwd(x, w, b):
    _cuda(x, w, b)
    M, K = x.shape
    Cc = w.shape[0]
    y = torch.empty(M, Cc, dtype=torch.float32, device=x.device)
    a = capi.HeadLinearFwdArgs(_p(x), _p(w), _p(b), _p(y), M, K, Cc, _dt(x))
    capi.call("tvt_head_linear_fwd", a, _stream())
    return y


def head_linear_bwd(x, w, dy, need_dx=True):
    M, K = x.shape
    Cc = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.zeros(Cc, K, dtype=torch.float32, device=x.device)
    db = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    a = capi.HeadLinearBwdArgs(_p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), M, K, Cc, _dt(x))
    capi.call("tvt_head_linear_bwd", a, _stream())
    return dx, dw, db


def cls_sum(tokens_list, B, S, d):
    _cuda(*tokens_list)
    out = torch.empty(B, d, dtype=tokens_list[0].dtype, device=tokens_list[0].device)
    a = capi.ClsSumArgs()
    for i, t in enumerate(tokens_list):
        a.tokens[i] = t.data_ptr()
    a.out, a.batch, a.seq_len, a.d, a.num_experts, a.dtype = _p(out), B, S, d, len(tokens_list), _dt(out)
    capi.call("tvt_cls_sum_fwd", a, _stream())
    return out


def distill_loss(student, teacher, target, *, w_bce=1.0, w_ce=0.0, w_kl=0.0, temperature=1.0, grad_scale=1.0,
                 need_grad=True):
    """student/teacher/target fp32 [B, C].  Returns (losses[5] = total,bce,ce,kl,cos0 ; dlogits or None)."""
    _cuda(student, teacher, target)
    B, Cc = student.shape
    losses = torch.zeros(5, dtype=torch.float32, device=student.device)
    dl = torch.empty_like(student) if need_grad else None
    a = capi.DistillLossArgs(_p(student), _p(teacher), _p(target), _p(losses), _p(dl), B, Cc, w_bce, w_ce, w_kl,
                             temperature, grad_scale)
    capi.call("tvt_distill_loss", a, _stream())
    return losses, dl


def pyramid_head(z, target=None, need_grad=False, grad_scale=1.0):
    """z fp32 [G, B, C] -> prob [B, C] (mean of sigmoids); optional BCE(prob, target) and dz."""
    _cuda(z, target)
    G, B, Cc = z.shape
    prob = torch.empty(B, Cc, dtype=torch.float32, device=z.device)
    loss = torch.zeros(1, dtype=torch.float32, device=z.device) if target is not None else None
    dz = torch.empty_like(z) if need_grad else None
    a = capi.PyramidHeadArgs(_p(z), _p(target), _p(prob), _p(loss), _p(dz), G, B, Cc, grad_scale)
    capi.call("tvt_pyramid_head", a, _stream())
    return prob, loss, dz


def eval_readout(logits, target, probs, labels, pred_bits, top1, row_offset, thresholds):
    """Evaluation read-out of one batch into caller-owned running buffers (tvt_eval_readout; reference:
    src/models/transformer.py:146-158 + src/callbacks/callbacks.py:34-45).  logits [B, C] fp32; target [B, C] fp32 /
    fp64 or None; probs / labels / pred_bits are [capacity, C], top1 is [capacity]."""
    _cuda(logits, probs)
    if logits.dtype != torch.float32 or not logits.is_contiguous():
        raise ValueError("eval_readout: logits must be contiguous fp32")
    if target is not None and target.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"eval_readout: target must be float32 or float64 (got {target.dtype})")
    if len(thresholds) > 16:
        raise ValueError("eval_readout: at most 16 thresholds")
    B, C = logits.shape
    a = capi.EvalReadoutArgs()
    a.logits, a.target, a.probs = _p(logits), _p(target.contiguous() if target is not None else None), _p(probs)
    a.labels, a.pred_bits, a.top1 = _p(labels), _p(pred_bits), _p(top1)
    a.batch, a.classes, a.row_offset, a.capacity = B, C, row_offset, probs.shape[0]
    for i, t in enumerate(thresholds):
        a.thresholds[i] = float(t)
    a.num_thresholds = len(thresholds)
    a.target_dtype = capi.TVT_F64 if (target is not None and target.dtype == torch.float64) else TVT_F32
    capi.call("tvt_eval_readout", a, _stream())


def feature_augment(x, d_out=None, *, p_drop=0.0, p_noise=0.0, noise_std=0.1 ** 0.5, seed=0, out_dtype=torch.bfloat16):
    """Loader-side pad + train-time feature drop / Gaussian noise + cast in one pass (tvt_feature_augment; reference:
    src/dataloaders/MMX_Temporal_dl.py:167-181).  x [..., d_in] fp32 -> [..., d_out] out_dtype."""
    _cuda(x)
    if x.dtype != torch.float32:
        raise ValueError("feature_augment: x must be float32 (the loader's dtype)")
    x = x.contiguous()
    d_in = x.shape[-1]
    d_out = d_in if d_out is None else int(d_out)
    y = torch.empty(*x.shape[:-1], d_out, dtype=out_dtype, device=x.device)
    a = capi.FeatureAugmentArgs(_p(x), _p(y), x.numel() // d_in, d_in, d_out, p_drop, p_noise, noise_std, seed, _dt(y))
    capi.call("tvt_feature_augment", a, _stream())
    return y


# ------------------------------------------------------------------------------------------ collaborative gating glue
def stretch_cast(x, out):
    """out[:, :] (rows of a [rows, 2048] slice, any row pitch) = nearest-neighbour stretch of x [rows, d_in] fp32."""
    _cuda(x, out)
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    a = capi.StretchCastArgs(_p(x), _p(out), x.shape[0], x.shape[1], out.shape[1], _rowmajor(out), _dt(out), 0)
    capi.call("tvt_stretch_cast", a, _stream())


def _collab(name, E, rows, d, dtype, **ptrs):
    a = capi.CollabArgs()
    for k, t in ptrs.items():
        setattr(a, k, _p(t))
    a.rows, a.d, a.experts, a.dtype = rows, d, E, dtype
    capi.call(name, a, _stream())


def collab_mix_fwd(c, pc):
    """c [E, N, D], pc [E-1, N, D] -> T [E, N, D] (see tvt_collab_mix_fwd)."""
    _cuda(c, pc)
    E, N, D = c.shape
    out = torch.empty_like(c)
    _collab("tvt_collab_mix_fwd", E, N, D, _dt(c), c=c, pc=pc, out=out)
    return out


def collab_mix_bwd(dt):
    E, N, D = dt.shape
    dc = torch.empty_like(dt)
    dpc = torch.empty(E - 1, N, D, dtype=dt.dtype, device=dt.device)
    _collab("tvt_collab_mix_bwd", E, N, D, _dt(dt), dout=dt, dc=dc, dpc=dpc)
    return dc, dpc


def collab_gate_fwd(c, a):
    _cuda(c, a)
    E, N, D = c.shape
    out = torch.empty(N, D, dtype=c.dtype, device=c.device)
    _collab("tvt_collab_gate_fwd", E, N, D, _dt(c), c=c, a=a, out=out)
    return out


def collab_gate_bwd(c, a, dg):
    E, N, D = c.shape
    dc, da = torch.empty_like(c), torch.empty_like(a)
    _collab("tvt_collab_gate_bwd", E, N, D, _dt(c), c=c, a=a, dout=dg, dc=dc, da=da)
    return dc, da


def l2norm_fwd(x, eps=1e-12):
    """x [rows, d] (bf16 / fp32) -> (y fp32 = x / max(||x||, eps), inv_norm [rows])."""
    _cuda(x)
    rows, d = x.shape
    y = torch.empty(rows, d, dtype=torch.float32, device=x.device)
    inv = torch.empty(rows, dtype=torch.float32, device=x.device)
    a = capi.L2NormArgs(_p(x), _p(y), _p(inv), None, None, rows, d, eps, _dt(x))
    capi.call("tvt_l2norm_fwd", a, _stream())
    return y, inv


def l2norm_bwd(dy, y, inv, dtype, eps=1e-12):
    rows, d = y.shape
    dx = torch.empty(rows, d, dtype=dtype, device=y.device)
    a = capi.L2NormArgs(None, _p(y), _p(inv), _p(dy), _p(dx), rows, d, eps, _dt(dx))
    capi.call("tvt_l2norm_bwd", a, _stream())
    return dx
