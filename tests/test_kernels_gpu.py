"""GPU: every C-ABI kernel against a plain fp32 torch restatement of the same op, on seeded inputs.
Tolerances: fp32 paths 1e-3 normwise relative (usually ~1e-6), bf16 paths 2e-2 (BASELINE.json north_star)."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tvt():
    import tvt_b200
    from tvt_b200 import ops
    assert tvt_b200.capi.load().tvt_device_check() == 0, tvt_b200.capi.last_error()
    return ops


def _dev():
    return torch.device("cuda:0")


def _bf(x):
    return x.to(torch.bfloat16)


# ----------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(136, 512, 2048), (2112, 1536, 512), (300, 136, 328), (8, 512, 16384), (128, 896, 128)])
def test_gemm_forward_bf16(tvt, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1130)
    x = _bf(torch.randn(M, K, device=_dev(), generator=g))
    w = _bf(torch.randn(N, K, device=_dev(), generator=g) / math.sqrt(K))
    b = torch.randn(N, device=_dev(), generator=g)
    r = _bf(torch.randn(M, N, device=_dev(), generator=g))
    out = torch.empty(M, N, dtype=torch.bfloat16, device=_dev())
    o32 = torch.empty(M, N, dtype=torch.float32, device=_dev())
    tvt.gemm(x, w, M, N, K, bias=b, act=tvt.ACT_RELU, residual=r, out_bf16=out, out_f32=o32)
    ref = F.relu(x.float() @ w.float().t() + b) + r.float()
    assert_close(o32, ref, 1e-5, "fp32 output")
    assert_close(out.float(), ref, 4e-3, "bf16 output")


@pytest.mark.parametrize("M,N,K", [(2112, 512, 2048), (264, 136, 200)])
def test_gemm_split_planes_fp32_accurate(tvt, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(M, K, device=_dev(), generator=g)
    w = torch.randn(N, K, device=_dev(), generator=g) / math.sqrt(K)
    xh, xl, wh, wl = (torch.empty(t.shape, dtype=torch.bfloat16, device=_dev()) for t in (x, x, w, w))
    tvt.split_f32(x, xh, xl)
    tvt.split_f32(w, wh, wl)
    assert_close(xh.float() + xl.float(), x, 1e-4, "hi+lo split")
    out = torch.empty(M, N, device=_dev())
    tvt.gemm(xh, wh, M, N, K, a_lo=xl, b_lo=wl, out_f32=out)
    ref = (x.double() @ w.double().t()).float()
    assert_close(out, ref, 5e-5, "3-pass GEMM")      # ~1e-5: two orders better than single-pass bf16


def test_gemm_dgrad_wgrad_layouts(tvt):
    M, N, K = 2112, 1536, 512       # tokens, out features, in features
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = _bf(torch.randn(M, N, device=_dev(), generator=g))
    w = _bf(torch.randn(N, K, device=_dev(), generator=g) / math.sqrt(K))
    x = _bf(torch.randn(M, K, device=_dev(), generator=g))
    dx = torch.empty(M, K, device=_dev())
    tvt.gemm(dy, w, M, K, N, b_mn=True, out_f32=dx)                       # dX = dY W
    assert_close(dx, dy.float() @ w.float(), 1e-5, "dgrad")
    dw = torch.zeros(N, K, device=_dev())
    tvt.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=dw, splits=4, atomic=True)   # dW = dY^T X
    assert_close(dw, dy.float().t() @ x.float(), 1e-5, "wgrad split-K")


@pytest.mark.parametrize("M,N,K,splits", [(33024, 3072, 768, 2), (33024, 2304, 768, 8), (2112, 1536, 512, 4), (4160, 512, 512, 1), (1000, 264, 520, 3)])
def test_gemm_wgrad_with_fused_bias_gradient(tvt, M, N, K, splits):
    """tvt_gemm_args.a_rowsum: colsum(dY) out of the weight-gradient GEMM itself (one extra N = 16 MMA per k-step against ones)
    = the bias gradient torch's Linear backward computes with a separate reduction (aten::sum over the token dimension)."""
    # 1e-4: the TMEM accumulator truncates (round toward zero) once per MMA, ~2e-5 over a 16 512-token chain with a non-zero mean
    g = torch.Generator(device="cuda").manual_seed(11)
    dy = _bf(torch.randn(M, N, device=_dev(), generator=g) + 0.25)
    x = _bf(torch.randn(M, K, device=_dev(), generator=g))
    if not tvt.rowsum_supported(N, K, M, splits):
        with pytest.raises(Exception, match="a_rowsum"):
            tvt.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=torch.zeros(N, K, device=_dev()), splits=splits, atomic=True,
                     a_rowsum=torch.zeros(N, device=_dev()))
        pytest.skip("kernel selection does not put this shape on the CTA-pair tiles: Mode.wgrad uses tvt_colsum for it")
    dw, db = torch.zeros(N, K, device=_dev()), torch.zeros(N, device=_dev())
    tvt.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=dw, splits=splits, atomic=True, a_rowsum=db)
    assert_close(dw, dy.float().t() @ x.float(), 1e-4, "wgrad with row sums")
    assert_close(db, dy.float().sum(0), 1e-4, "bias gradient from the wgrad GEMM")
    # accumulating launch (DDP gradient sinks): both outputs add
    tvt.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=dw, splits=splits, atomic=True, a_rowsum=db)
    assert_close(db, 2 * dy.float().sum(0), 1e-4, "bias gradient accumulates")
    assert_close(dw, 2 * (dy.float().t() @ x.float()), 1e-4, "wgrad accumulates")
    # Mode.wgrad picks the fused path by itself and agrees with the column-sum kernel
    m = tvt.Mode("bf16")
    db2, db3 = torch.zeros(N, device=_dev()), torch.zeros(N, device=_dev())
    m.wgrad((dy, None), (x, None), M, N, K, bias_grad=db2, dy=dy)
    tvt.colsum(dy, db3)
    assert_close(db2, db3, 1e-4, "Mode.wgrad bias gradient vs tvt_colsum")


def test_gemm_fused_backward_epilogue(tvt):
    """dgrad with ReLU mask + dropout regeneration + residual: the linear2 -> linear1 hop of the FFN backward."""
    M, N, K = 512, 256, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = _bf(torch.randn(M, N, device=_dev(), generator=g))
    w = _bf(torch.randn(N, K, device=_dev(), generator=g) / 16)
    h = _bf(torch.relu(torch.randn(M, K, device=_dev(), generator=g)))
    out = torch.empty(M, K, device=_dev())
    tvt.gemm(dy, w, M, K, N, b_mn=True, relu_mask=h, out_f32=out)
    ref = (dy.float() @ w.float()) * (h.float() > 0)
    assert_close(out, ref, 1e-5, "relu-masked dgrad")
    # forward dropout mask == mask regenerated by a second launch with the same seed
    x = _bf(torch.randn(M, N, device=_dev(), generator=g))
    wf = _bf(torch.randn(K, N, device=_dev(), generator=g) / 16)
    y1 = torch.empty(M, K, device=_dev()); y2 = torch.empty(M, K, device=_dev()); y0 = torch.empty(M, K, device=_dev())
    tvt.gemm(x, wf, M, K, N, out_f32=y0)
    tvt.gemm(x, wf, M, K, N, out_f32=y1, dropout_p=0.5, seed=99)
    tvt.gemm(x, wf, M, K, N, out_f32=y2, dropout_p=0.5, seed=99)
    assert torch.equal(y1, y2)
    kept = y1 != 0
    assert 0.47 < kept.float().mean().item() < 0.53
    assert_close(y1[kept], 2.0 * y0[kept], 1e-6, "dropout scale")


def test_gemm_rejects_bad_arguments(tvt):
    from tvt_b200 import TvtError
    x = torch.zeros(16, 64, dtype=torch.bfloat16, device=_dev())
    w = torch.zeros(15, 64, dtype=torch.bfloat16, device=_dev())
    out = torch.zeros(16, 16, device=_dev())
    with pytest.raises(TvtError, match="multiple of 8"):
        tvt.gemm(x, w, 16, 15, 64, out_f32=out)
    with pytest.raises(TvtError, match="CUDA"):
        tvt.gemm(x.cpu(), w, 16, 8, 64, out_f32=out)


# ----------------------------------------------------------------------------------------- LayerNorm / embed
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("rows,d", [(136, 512), (2112, 768), (42, 896), (17, 2048), (5, 192)])
def test_layernorm_fwd_bwd(tvt, dtype, tol, rows, d):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.randn(rows, d, device=_dev(), generator=g) * 2 + 0.5).to(dtype)
    gamma = torch.randn(d, device=_dev(), generator=g)
    beta = torch.randn(d, device=_dev(), generator=g)
    dy = torch.randn(rows, d, device=_dev(), generator=g).to(dtype)
    y, mean, rstd = tvt.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), gr, br, 1e-5)
    yr.backward(dy.float())
    assert_close(y.float(), yr, tol, "ln fwd")
    dg, db, dbias = (torch.zeros(d, device=_dev()) for _ in range(3))
    dx, dz = tvt.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dbias=dbias)
    assert dz is dx
    assert_close(dx.float(), xr.grad, tol, "ln dx")
    assert_close(dg, gr.grad, max(tol, 1e-4), "ln dgamma")
    assert_close(db, br.grad, max(tol, 1e-4), "ln dbeta")
    assert_close(dbias, dx.float().sum(0), 1e-3 if dtype == torch.float32 else 2e-2, "ln dbias = colsum(dx)")


def test_layernorm_full_size_properties(tvt):
    """C5 size (one more row than 256 clips x 129 tokens, so the last ring slot holds a single row): size-independent
    properties of LayerNorm and its backward on every row, plus torch on a strided sample of rows."""
    rows, d = 256 * 129 + 1, 768
    g = torch.Generator(device="cuda").manual_seed(17)
    x = (torch.randn(rows, d, device=_dev(), generator=g) * 1.5 + 0.25).to(torch.bfloat16)
    dy = torch.randn(rows, d, device=_dev(), generator=g).to(torch.bfloat16)
    gamma = 1.0 + 0.1 * torch.randn(d, device=_dev(), generator=g)
    beta = 0.1 * torch.randn(d, device=_dev(), generator=g)
    ones, zeros = torch.ones(d, device=_dev()), torch.zeros(d, device=_dev())
    y, mean, rstd = tvt.layernorm_fwd(x, ones, zeros)
    yf = y.float()
    assert float(yf.mean(1).abs().max()) < 2e-2 and float((yf.var(1, unbiased=False) - 1).abs().max()) < 3e-2   # bf16 output rounding
    assert_close(mean, x.float().mean(1), 1e-5, "row means")
    assert_close(rstd, (x.float().var(1, unbiased=False) + 1e-5).rsqrt(), 1e-5, "row rstd")
    y2, _, _ = tvt.layernorm_fwd(x, gamma, beta)
    dg, db, dbias = (torch.zeros(d, device=_dev()) for _ in range(3))
    res = torch.randn(rows, d, device=_dev(), generator=g).to(torch.bfloat16)
    dx, _ = tvt.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=dg, dbeta=db, dbias=dbias)
    dxr, _ = tvt.layernorm_bwd(dy, x, mean, rstd, gamma, dres=res)
    # backward invariants: dx is orthogonal to the constant vector and to x_hat, row by row
    xhat = (x.float() - mean[:, None]) * rstd[:, None]
    scale = (dy.float() * gamma).abs().sum(1) * rstd
    assert float((dx.float().sum(1).abs() / scale).max()) < 1e-2
    assert float(((dx.float() * xhat).sum(1).abs() / (scale * math.sqrt(d))).max()) < 1e-2
    assert_close(dxr.float(), dx.float() + res.float(), 8e-3, "residual-path gradient is added")
    assert_close(db, dy.float().sum(0), 1e-4, "dbeta = colsum(dy)")
    assert_close(dg, (dy.float() * xhat).sum(0), 1e-4, "dgamma = colsum(dy * xhat)")
    assert_close(dbias, dx.float().sum(0), 2e-2, "dbias = colsum(dx)")
    idx = torch.arange(0, rows, 257, device=_dev())
    idx = torch.cat((idx, torch.tensor([rows - 2, rows - 1], device=_dev())))
    xr = x[idx].float().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), gamma, beta, 1e-5)
    yr.backward(dy[idx].float())
    assert_close(y2[idx].float(), yr, 1e-2, "sampled rows fwd")
    assert_close(dx[idx].float(), xr.grad, 1e-2, "sampled rows dx")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(200, 256), (130, 768), (33, 896)])
def test_layernorm_bwd_regenerates_gemm_dropout_mask(tvt, dtype, rows, d):
    """The branch gradient dz of the LayerNorm backward must carry exactly the mask the producing GEMM's epilogue
    applied in forward (same seed, element index = row * d + col), scaled by 1 / (1 - p)."""
    g = torch.Generator(device="cuda").manual_seed(13)
    a = _bf(torch.randn(rows, 64, device=_dev(), generator=g))
    w = _bf(torch.randn(d, 64, device=_dev(), generator=g))
    y0 = torch.empty(rows, d, device=_dev()); y1 = torch.empty(rows, d, device=_dev())
    tvt.gemm(a, w, rows, d, 64, out_f32=y0)
    tvt.gemm(a, w, rows, d, 64, out_f32=y1, dropout_p=0.5, seed=4242)
    kept = (y1 != 0) | (y0 == 0)
    x = (torch.randn(rows, d, device=_dev(), generator=g)).to(dtype)
    dy = torch.randn(rows, d, device=_dev(), generator=g).to(dtype)
    gamma = torch.randn(d, device=_dev(), generator=g)
    _, mean, rstd = tvt.layernorm_fwd(x, gamma, torch.zeros_like(gamma))
    dbias = torch.zeros(d, device=_dev())
    dx, dz = tvt.layernorm_bwd(dy, x, mean, rstd, gamma, dbias=dbias, dropout_p=0.5, seed=4242)
    dx0, _ = tvt.layernorm_bwd(dy, x, mean, rstd, gamma)
    assert_close(dx.float(), dx0.float(), 1e-6, "dx is not masked")
    ref = torch.where(kept, 2.0 * dx0.float(), torch.zeros_like(dx0.float()))
    assert_close(dz.float(), ref, 1e-6 if dtype == torch.float32 else 8e-3, "dz == forward mask * dx / (1 - p)")
    assert torch.equal(dz != 0, kept & (dx0 != 0))
    assert_close(dbias, dz.float().sum(0), 1e-3 if dtype == torch.float32 else 2e-2, "dbias = colsum(dz)")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_embed_prologue(tvt, dtype, tol):
    """add_pos_cls (transformer.py:74-82): CLS concat + PE + LN, and its backward scatter."""
    from oracle import param
    B, T, d = 6, 16, 512
    g = torch.Generator(device="cuda").manual_seed(2)
    feat = torch.randn(B, T, d, device=_dev(), generator=g).to(dtype)
    cls = torch.rand(B, d, device=_dev(), generator=g).to(dtype)
    gamma = torch.randn(d, device=_dev(), generator=g); beta = torch.randn(d, device=_dev(), generator=g)
    pe = param.PositionalEncoding(d, 0.0, max_len=T + 1).pe[:, 0, :].to(_dev()).contiguous()
    y, pre, mean, rstd = tvt.embed_fwd(feat, cls, pe, gamma, beta)
    fr, cr = feat.float().requires_grad_(True), cls.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    tok = torch.cat((cr.unsqueeze(1), fr), dim=1) + pe.unsqueeze(0)
    ref = F.layer_norm(tok, (d,), gr, br, 1e-5)
    assert_close(y.float().view(B, T + 1, d), ref, tol, "embed fwd")
    dy = torch.randn(B * (T + 1), d, device=_dev(), generator=g).to(dtype)
    ref.backward(dy.float().view(B, T + 1, d))
    dg, db = torch.zeros(d, device=_dev()), torch.zeros(d, device=_dev())
    dfeat, dcls = tvt.embed_bwd(dy, pre, mean, rstd, gamma, B, T + 1, dgamma=dg, dbeta=db)
    assert_close(dfeat.float(), fr.grad, tol, "embed dfeat")
    assert_close(dcls.float(), cr.grad, tol, "embed dcls")
    assert_close(dg, gr.grad, max(tol, 1e-4), "embed dgamma")


# ----------------------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, B, H, Sq, Sk, hd, scale):
    qh = q.view(B, Sq, H, hd).transpose(1, 2)
    kh = k.view(B, Sk, H, hd).transpose(1, 2)
    vh = v.view(B, Sk, H, hd).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B * Sq, H * hd)


@pytest.mark.parametrize("impl,dtype,tol", [(1, torch.float32, 1e-5), (1, torch.bfloat16, 1.5e-2), (2, torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("B,H,Sq,Sk,hd", [(8, 8, 17, 17, 64), (4, 8, 33, 66, 64), (3, 2, 14, 14, 448), (2, 12, 129, 129, 64), (2, 3, 5, 5, 64),
                                          (3, 4, 129, 258, 64), (2, 2, 65, 65, 64), (1, 1, 256, 272, 64), (5, 2, 1, 1, 64),
                                          (2, 2, 136, 136, 64), (2, 3, 131, 144, 64), (2, 2, 137, 129, 64), (2, 2, 128, 128, 64)])
def test_attention_fwd_bwd(tvt, impl, dtype, tol, B, H, Sq, Sk, hd):
    if impl == 2 and hd != 64:
        pytest.skip("tcgen05 attention kernel is head_dim 64 only")
    if impl == 1 and Sk > 258:
        pytest.skip("beyond the shared-memory resident CUDA-core kernel")
    g = torch.Generator(device="cuda").manual_seed(1130)
    d = H * hd
    # packed projections, as the in-proj GEMM writes them: q | k | v side by side (self) or q, kv (cross)
    qbuf = torch.randn(B * Sq, 3 * d, device=_dev(), generator=g).to(dtype)
    kvbuf = qbuf if Sq == Sk else torch.randn(B * Sk, 3 * d, device=_dev(), generator=g).to(dtype)
    q, k, v = qbuf[:, :d], kvbuf[:, d:2 * d], kvbuf[:, 2 * d:]
    scale = 1.0 / math.sqrt(hd)
    o, lse = tvt.attention_fwd(q, k, v, B, H, Sq, Sk, hd, scale, impl=impl)
    qr, kr, vr = (t.float().contiguous().requires_grad_(True) for t in (q, k, v))
    ref = _sdpa_ref(qr, kr, vr, B, H, Sq, Sk, hd, scale)
    assert_close(o.float(), ref, tol, "attention fwd")
    do = torch.randn(B * Sq, d, device=_dev(), generator=g).to(dtype)
    ref.backward(do.float())
    dq, dk, dv = (torch.empty(t.shape, dtype=dtype, device=_dev()) for t in (qr, kr, vr))
    tvt.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Sq, Sk, hd, scale, impl=impl)
    assert_close(dq.float(), qr.grad, tol, "attention dq")
    assert_close(dk.float(), kr.grad, tol, "attention dk")
    assert_close(dv.float(), vr.grad, tol, "attention dv")


def test_attention_impls_share_dropout_masks(tvt):
    """The tcgen05 forward and the CUDA-core backward (and vice versa) must regenerate the same mask
    (including the CUDA-core tail rows / tail keys of the 2^k + 1 token sequences)."""
    hd = 64
    for B, H, Sq, Sk in ((3, 4, 33, 33), (2, 3, 129, 129), (2, 2, 129, 258), (2, 2, 140, 140), (2, 2, 136, 144), (3, 2, 130, 130)):
        g = torch.Generator(device="cuda").manual_seed(21)
        q = torch.randn(B * Sq, H * hd, device=_dev(), generator=g).to(torch.bfloat16)
        k, v = (torch.randn(B * Sk, H * hd, device=_dev(), generator=g).to(torch.bfloat16) for _ in range(2))
        o1, l1 = tvt.attention_fwd(q, k, v, B, H, Sq, Sk, hd, 0.125, dropout_p=0.5, seed=77, impl=1)
        o2, l2 = tvt.attention_fwd(q, k, v, B, H, Sq, Sk, hd, 0.125, dropout_p=0.5, seed=77, impl=2)
        assert_close(o2.float(), o1.float(), 1.5e-2, f"dropout fwd simt vs tcgen05 {Sq}x{Sk}")
        assert_close(l2, l1, 1e-4, "lse")
        do = torch.randn(B * Sq, H * hd, device=_dev(), generator=g).to(torch.bfloat16)
        grads = []
        for impl in (1, 2):
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            tvt.attention_bwd(q, k, v, o1, do, l1, dq, dk, dv, B, H, Sq, Sk, hd, 0.125, dropout_p=0.5, seed=77, impl=impl)
            grads.append((dq, dk, dv))
        for a, b, n in zip(grads[0], grads[1], "qkv"):
            assert_close(b.float(), a.float(), 2e-2, f"dropout d{n} simt vs tcgen05 {Sq}x{Sk}")


def test_attention_dropout_statistics(tvt):
    B, H, S, hd = 4, 4, 33, 64
    g = torch.Generator(device="cuda").manual_seed(4)
    q = torch.zeros(B * S, H * hd, device=_dev())              # uniform attention: every prob = 1/S
    v = torch.ones(B * S, H * hd, device=_dev())
    o0, _ = tvt.attention_fwd(q, q, v, B, H, S, S, hd, 1.0)
    o1, _ = tvt.attention_fwd(q, q, v, B, H, S, S, hd, 1.0, dropout_p=0.5, seed=123)
    o2, _ = tvt.attention_fwd(q, q, v, B, H, S, S, hd, 1.0, dropout_p=0.5, seed=123)
    assert torch.equal(o1, o2)
    assert_close(o0, torch.ones_like(o0), 1e-5, "no-dropout")
    # each output = (kept count / S) * 2: mean 1, variance 1/S
    assert abs(o1[:, ::hd].mean().item() - 1.0) < 0.05
    assert 0.5 / S < o1[:, ::hd].var().item() < 2.0 / S


# ----------------------------------------------------------------------------------------- pooling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,d", [(5, 20, 64), (4, 64, 512), (3, 9, 32)])
def test_pyramid_pool(tvt, dtype, B, T, d):
    from oracle import param
    g = torch.Generator(device="cuda").manual_seed(8)
    S = T + 1
    tok = torch.randn(B * S, d, device=_dev(), generator=g).to(dtype)
    groups = [2, 3, 4]
    outs = tvt.pyramid_pool_fwd(tok, B, S, d, groups, relu=True)
    frames = tok.float().view(B, S, d)[:, 1:, :].detach().requires_grad_(True)
    refs = [torch.relu(param.sum_group(frames, gp)) for gp in groups]
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    for o, r, gp in zip(outs, refs, groups):
        assert o.shape == (B, (T // gp) * d)
        assert_close(o.float(), r, tol, f"sum_group g={gp}")
    douts = [torch.randn(o.shape, device=_dev(), generator=g).to(dtype) for o in outs]
    # reference backward uses OUR forward outputs for the ReLU mask decision boundary (identical in fp32)
    sum((torch.relu(param.sum_group(frames, gp)) * do.float()).sum() for gp, do in zip(groups, douts)).backward()
    dtok = torch.full((B * S, d), 7.0, device=_dev()).to(dtype)
    tvt.pyramid_pool_bwd(douts, outs, dtok, B, S, d, groups, relu=True)
    got = dtok.float().view(B, S, d)
    assert torch.all(got[:, 0] == 7.0)                       # CLS rows untouched
    assert_close(got[:, 1:], frames.grad, 1e-6 if dtype == torch.float32 else 2e-2, "pool bwd")


def test_spatial_pool(tvt):
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(6, 128, 28, 28, device=_dev(), generator=g)
    out = torch.zeros(6, 896, device=_dev())
    tvt.spatial_pool(x, out, 768)
    assert_close(out[:, 768:], x.mean(dim=(2, 3)), 1e-6, "spatial pool")
    assert float(out[:, :768].abs().max()) == 0.0


# ----------------------------------------------------------------------------------------- heads and losses
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_head_linear(tvt, dtype, tol):
    M, K, C = 64, 512, 15
    g = torch.Generator(device="cuda").manual_seed(10)
    x = torch.randn(M, K, device=_dev(), generator=g).to(dtype)
    w = torch.randn(C, K, device=_dev(), generator=g) / 20
    b = torch.randn(C, device=_dev(), generator=g)
    y = tvt.head_linear_fwd(x, w, b)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.linear(xr, wr, br)
    assert_close(y, ref, 1e-5, "head fwd")
    dy = torch.randn(M, C, device=_dev(), generator=g)
    ref.backward(dy)
    dx, dw, db = tvt.head_linear_bwd(x, w, dy)
    assert_close(dx.float(), xr.grad, tol, "head dx")
    assert_close(dw, wr.grad, 1e-5, "head dw")
    assert_close(db, br.grad, 1e-5, "head db")


@pytest.mark.parametrize("B,C", [(8, 15), (256, 15), (3, 19), (5, 40)])
def test_distill_loss(tvt, B, C):
    from oracle import param
    g = torch.Generator(device="cuda").manual_seed(12)
    s = torch.randn(B, C, device=_dev(), generator=g) * 2
    t = torch.randn(B, C, device=_dev(), generator=g) * 2
    y = (torch.rand(B, C, device=_dev(), generator=g) < 0.15).float()
    T, alpha = 2.0, 0.7
    losses, dl = tvt.distill_loss(s, t, y, w_bce=1.0, w_ce=1.0, w_kl=alpha, temperature=T)
    sr = s.clone().requires_grad_(True)
    ref, parts = param.distill_loss(sr, t, y, temperature=T, alpha=alpha)
    ref.backward()
    assert_close(losses[0], ref, 1e-5, "total")
    assert_close(losses[1], parts["base"], 1e-5, "bce")
    assert_close(losses[2], parts["distil"], 1e-5, "ce")
    assert_close(losses[3], parts["kl"], 1e-4, "kl")
    assert_close(losses[4], parts["cos"], 1e-5, "cos")
    assert_close(dl, sr.grad, 1e-5, "dlogits")
    # reference-only form (frame_transformer.py:250-252): BCE + hard-label CE, and plain BCE (transformer.py:142)
    l2, _ = tvt.distill_loss(s, t, y, w_bce=1.0, w_ce=1.0)
    assert_close(l2[0], param.distill_loss(s, t, y)[0], 1e-5, "bce+ce")
    l3, _ = tvt.distill_loss(s, None, y)
    assert_close(l3[0], F.binary_cross_entropy_with_logits(s, y), 1e-5, "bce only")


def test_pyramid_head(tvt):
    G, B, C = 3, 16, 15
    g = torch.Generator(device="cuda").manual_seed(13)
    z = torch.randn(G, B, C, device=_dev(), generator=g)
    y = (torch.rand(B, C, device=_dev(), generator=g) < 0.15).float()
    prob, loss, dz = tvt.pyramid_head(z, y, need_grad=True)
    zr = z.clone().requires_grad_(True)
    pr = torch.sigmoid(zr).mean(0)
    lr = F.binary_cross_entropy(pr, y)
    lr.backward()
    assert_close(prob, pr, 1e-6, "prob")
    assert_close(loss[0], lr, 1e-5, "bce")
    assert_close(dz, zr.grad, 1e-4, "dz")


def test_colsum_bias_act_actbwd(tvt):
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.randn(1000, 776, device=_dev(), generator=g)
    out = torch.zeros(776, device=_dev())
    tvt.colsum(x, out)
    assert_close(out, x.sum(0), 1e-5, "colsum f32")
    xb = x.to(torch.bfloat16)
    out.zero_()
    tvt.colsum(xb, out)
    assert_close(out, xb.float().sum(0), 1e-5, "colsum bf16")
    bias = torch.randn(776, device=_dev(), generator=g)
    y = torch.empty(1000, 776, device=_dev())
    tvt.bias_act(x, bias, y, tvt.ACT_GELU)
    assert_close(y, F.gelu(x + bias), 1e-5, "bias+gelu")
    dy = torch.randn(1000, 776, device=_dev(), generator=g)
    xr = x.clone().requires_grad_(True)
    F.gelu(xr).backward(dy)
    assert_close(tvt.act_bwd(dy, x, tvt.ACT_GELU), xr.grad, 1e-5, "gelu bwd")
    assert_close(tvt.act_bwd(dy, x, tvt.ACT_RELU), dy * (x > 0), 1e-6, "relu bwd")


@pytest.mark.parametrize("kind", ["adamw", "sgd", "adagrad"])
def test_flat_optimizer_matches_torch(tvt, kind):
    """the flat-bucket step equals torch.optim.AdamW / SGD(momentum, weight_decay) over
    several steps, keeps parameters as views of the flat buffer, and refreshes the bf16 weight planes."""
    from tvt_b200 import ddp, optim
    torch.manual_seed(5)
    lin = torch.nn.Sequential(torch.nn.Linear(96, 130), torch.nn.LayerNorm(130), torch.nn.Linear(130, 7)).to(_dev())
    ref = torch.nn.Sequential(torch.nn.Linear(96, 130), torch.nn.LayerNorm(130), torch.nn.Linear(130, 7)).to(_dev())
    ref.load_state_dict(lin.state_dict())
    mode = tvt.Mode("bf16")
    red = ddp.GradBucketReducer(list(lin.parameters()), bucket_bytes=4096)
    assert len(red.buckets) >= 2
    if kind == "adamw":
        mine = optim.FlatOptimizer(red, modes=[mode], kind="adamw", lr=3e-3, weight_decay=0.09)
        theirs = torch.optim.AdamW(ref.parameters(), lr=3e-3, weight_decay=0.09)
    elif kind == "adagrad":
        mine = optim.FlatOptimizer(red, modes=[mode], kind="adagrad", lr=3e-2, weight_decay=0.09, eps=1e-10)
        theirs = torch.optim.Adagrad(ref.parameters(), lr=3e-2, weight_decay=0.09)
    else:
        mine = optim.FlatOptimizer(red, modes=[mode], kind="sgd", lr=3e-2, weight_decay=0.09, momentum=0.5)
        theirs = torch.optim.SGD(ref.parameters(), lr=3e-2, weight_decay=0.09, momentum=0.5)
    g = torch.Generator(device="cuda").manual_seed(6)
    for _ in range(4):
        x = torch.randn(32, 96, device=_dev(), generator=g)
        mine.zero_grad()
        theirs.zero_grad()
        lin(x).square().mean().backward()
        ref(x).square().mean().backward()
        red.finish()
        mine.step()
        theirs.step()
    for (n, a), (_, b) in zip(lin.named_parameters(), ref.named_parameters()):
        assert_close(a, b, 2e-6, f"{kind} {n}")
    w = lin[0].weight
    hi, lo = mode.weight(w)
    assert lo is None and torch.equal(hi, w.detach().to(torch.bfloat16))      # planes refreshed by the step kernel


@pytest.mark.parametrize("d_in,d_out,dtype", [(128, 2048, "bf16"), (1024, 1024, "f32"), (2048, 2048, "bf16"), (8, 16, "f32")])
def test_feature_augment_matches_oracle(tvt, d_in, d_out, dtype):
    """Loader pad + feature drop + Gaussian noise + cast (MMX_Temporal_dl.py:167-181) against the numpy oracle driven
    by the same counter-based stream: every drop / noise decision identical, values to float rounding; and the
    distribution the reference asks for (0.3 / 0.3, variance 0.1, independent decisions)."""
    import numpy as np
    from oracle import augment
    rows = 4099 if d_out > 100 else 60000
    g = torch.Generator().manual_seed(11)
    x = torch.randn(rows, d_in, generator=g)
    seed = 0x1234_5678_9ABC_DEF1
    odt = torch.bfloat16 if dtype == "bf16" else torch.float32
    y = tvt.feature_augment(x.to(_dev()), d_out, p_drop=0.3, p_noise=0.3, seed=seed, out_dtype=odt)
    want, drop, noisy = augment.feature_augment(x.numpy(), d_out, 0.3, 0.3, 0.1 ** 0.5, seed)
    got = y.float().cpu().numpy()
    assert got.shape == (rows, d_out)
    # decisions, read back from the output: dropped and noise-free rows are exactly zero everywhere; the padded columns of
    # a noise-free row are exactly zero
    zero_rows = ~got.any(axis=1)
    assert np.array_equal(zero_rows, drop & ~noisy)
    if d_out > d_in:
        assert np.array_equal(got[:, d_in:].any(axis=1), noisy)
    tol = 1e-5 if dtype == "f32" else 2.0 ** -7
    assert np.all(np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want)))
    if rows >= 60000:
        assert abs(drop.mean() - 0.3) < 0.01 and abs(noisy.mean() - 0.3) < 0.01
        assert abs((drop & noisy).mean() - 0.09) < 0.006                       # independent decisions
        z = got[drop & noisy]                                                   # pure noise rows
        assert abs(z.var() - 0.1) < 0.003 and abs(z.mean()) < 0.003
        assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03                  # the Box-Muller pair is uncorrelated
    # evaluation state: pad + cast only
    y0 = tvt.feature_augment(x.to(_dev()), d_out, p_drop=0.0, p_noise=0.0, seed=seed, out_dtype=odt).float().cpu()
    ref0 = torch.nn.functional.pad(x, (0, d_out - d_in))
    assert torch.equal(y0, ref0.to(odt).float())
    with pytest.raises(ValueError):
        tvt.feature_augment(x.double().to(_dev()), d_out)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 6e-3)])
def test_collaborative_gating_glue_kernels(dtype, tol):
    """csrc/collab.cu against torch: the nearest-neighbour stretch (F.interpolate), the pairwise-sum bookkeeping T, the GLU
    gate sum and the L2 normalisation, each forward and backward (collabgating.py:12-16,35-53,66-69,83-85)."""
    import torch.nn.functional as F
    from tvt_b200 import ops
    import tvt_b200  # noqa: F401
    from tvt_b200.functions import CollabGateFn, CollabMixFn, L2NormFn
    DEV = _dev()
    gen = torch.Generator().manual_seed(11)
    E, N, D = 3, 37, 256
    # stretch + cast into a slice of a stacked operand
    for d_in in (128, 200, 256):
        x = torch.randn(N, d_in, generator=gen).to(DEV)
        X = torch.zeros(2 * N, D, dtype=dtype, device=DEV)
        ops.stretch_cast(x, X[N:])
        want = F.interpolate(x.unsqueeze(0), D).squeeze(0)
        assert_close(X[N:].float(), want.to(dtype).float(), 0.0 if dtype == torch.float32 else 1e-7, f"stretch {d_in}")
        assert float(X[:N].abs().max()) == 0.0
    c = torch.randn(E, N, D, generator=gen).to(DEV).to(dtype).requires_grad_(True)
    pc = torch.randn(E - 1, N, D, generator=gen).to(DEV).to(dtype).requires_grad_(True)
    a = torch.randn(E, N, D, generator=gen).to(DEV).to(dtype).requires_grad_(True)
    w = torch.randn(E, N, D, generator=gen).to(DEV)
    cr, pcr, ar = (t.detach().float().requires_grad_(True) for t in (c, pc, a))
    # mix: the reference's bookkeeping written with cumulative sums (hostapi docstring)
    suffix = torch.flip(torch.cumsum(torch.flip(cr, [0]), 0), [0])
    T_ref = (E - 1) * cr + (suffix - cr)
    T_ref = torch.cat((T_ref[:1], T_ref[1:] + torch.cumsum(pcr, 0)))
    (T_ref * w).sum().backward()
    T = CollabMixFn.apply(c, pc)
    (T.float() * w).sum().backward()
    assert_close(T, T_ref, tol, "mix fwd")
    assert_close(c.grad, cr.grad, tol, "mix dC")
    assert_close(pc.grad, pcr.grad, tol, "mix dPC")
    # gate
    c.grad = None; cr.grad = None
    g_ref = (cr * torch.sigmoid(cr + ar)).sum(0)
    (g_ref * w[0]).sum().backward()
    g = CollabGateFn.apply(c, a)
    (g.float() * w[0]).sum().backward()
    assert_close(g, g_ref, tol, "gate fwd")
    assert_close(c.grad, cr.grad, tol, "gate dC")
    assert_close(a.grad, ar.grad, tol, "gate dA")
    # l2 normalisation, including a zero row (norm clamped by eps)
    x = torch.randn(N, D, generator=gen)
    x[5] = 0.0
    x = x.to(DEV).to(dtype).requires_grad_(True)
    xr = x.detach().float().requires_grad_(True)
    y_ref = F.normalize(xr)
    (y_ref * w[1]).sum().backward()
    y = L2NormFn.apply(x)
    (y * w[1]).sum().backward()
    assert y.dtype == torch.float32
    assert_close(y, y_ref, 1e-6, "l2norm fwd")
    keep = torch.ones(N, dtype=torch.bool); keep[5] = False      # the clamped row's gradient is dy / eps: compare the others
    assert_close(x.grad[keep.to(DEV)], xr.grad[keep.to(DEV)], tol, "l2norm dx")
    assert torch.isfinite(x.grad.float()).all()
