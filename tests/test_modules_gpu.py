"""GPU: the reference-facing modules (tvt_b200.hostapi) against the CPU-restated oracle (oracle/param.py)
run in fp32 on the same device, weights copied through state_dict, identical seeded inputs, dropout 0.
Bars (BASELINE.json north_star): fp32-accumulate mode 1e-3, bf16 mode 2e-2 (normwise relative) on logits
and on every parameter gradient (the top-1 criterion on 10 000 clips and the BASELINE-size cases live in
tests/test_fullsize_gpu.py)."""
import os

import pytest
import torch

from util import assert_close, copy_state, grads_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def api():
    import tvt_b200
    from tvt_b200 import hostapi
    assert tvt_b200.capi.load().tvt_device_check() == 0
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return hostapi


def _no_dropout(*mods):
    """Reasoning hard-codes Dropout(0.6) / Dropout(0.5) (TPN.py:92,95): parity runs zero every dropout so the
    train-mode oracle is deterministic (dropout itself is checked statistically elsewhere)."""
    for m in mods:
        for sub in m.modules():
            if isinstance(sub, torch.nn.Dropout):
                sub.p = 0.0
    return mods[0] if len(mods) == 1 else mods


def _yardstick(ref, precision, run):
    """Gradients of a copy of the fp32 oracle under stock PyTorch's matching reduced-precision mode (util.
    stock_reduced_precision: bf16 autocast / TF32 matmuls) — what plain PyTorch gives on the same weights and inputs.
    `run(module)` must return the scalar loss and wrap its forward in `_ac`."""
    import copy
    import util
    y = copy.deepcopy(ref)
    y.zero_grad(set_to_none=True)
    util.YARD_PRECISION[0] = precision
    run(y).float().backward()
    return y


def _ac(fn):
    """Run a forward under the stock reduced-precision mode of the yardstick (losses are evaluated outside, in fp32)."""
    import util
    return util.reduced(fn)


def _targets(B, C, gen):
    y = (torch.rand(B, C, generator=gen) < 0.15).float()
    y[torch.arange(B), torch.randint(0, C, (B,), generator=gen)] = 1.0       # loader forces one positive
    return y


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_simple_transformer_ptn_parity(api, precision):
    from oracle import param
    B = 4 if precision == "fp32" else 128       # bf16: enough tokens that ReLU-gate flips average out (tools/diag_bf16.py)
    cfg = dict(batch_size=B, seq_len=16, cls=1, dropout=0.0, input_dimension=256, nhead=4, nhid=512, nlayers=2,
               model="ptn", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15)
    torch.manual_seed(1130)
    ref = param.SimpleTransformer(**cfg).to(DEV)
    mod = copy_state(api.SimpleTransformer(precision=precision, **cfg), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    x = torch.randn(B, 16, 3, 256, generator=gen).to(DEV)                   # third expert bypasses the encoders
    y = _targets(B, 15, gen).to(DEV)
    lr = ref.criterion(ref.ptn(x), y)
    lr.backward()
    loss = mod.training_step({"experts": x, "label": y}, 0)
    loss.backward()
    assert_close(mod.ptn(x), ref.ptn(x), TOL[precision], "logits")
    assert_close(loss, lr, TOL[precision], "loss")
    yard = _yardstick(ref, precision, lambda m: m.criterion(_ac(lambda: m.ptn(x)).float(), y))
    worst = grads_close(mod, ref, TOL[precision], "ptn ", skip=("mlp_encoder", "encoder_layers"), yard=yard)
    print("worst grad", precision, worst)



@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_simple_transformer_ptn_shared_parity(api, precision):
    """Hierarchical fusion (src/models/transformer.py:84-104, intent restored as in oracle.param): ONE encoder applied to
    every expert, then a second encoder over the E expert CLS tokens (S = E + 1 = 4)."""
    from oracle import param
    B = 4 if precision == "fp32" else 128
    cfg = dict(batch_size=B, seq_len=16, cls=1, dropout=0.0, input_dimension=256, nhead=4, nhid=512, nlayers=2,
               model="ptn_shared", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15)
    torch.manual_seed(1130)
    ref = param.SimpleTransformer(**cfg).to(DEV)
    mod = copy_state(api.SimpleTransformer(precision=precision, **cfg), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    x = torch.randn(B, 16, 3, 256, generator=gen).to(DEV)
    y = _targets(B, 15, gen).to(DEV)
    lr = ref.criterion(ref.ptn_shared(x), y)
    lr.backward()
    logits = mod.ptn_shared(x)
    loss = mod._loss(logits, y)
    loss.backward()
    assert_close(logits, ref.ptn_shared(x), TOL[precision], "logits")
    assert_close(loss, lr, TOL[precision], "loss")
    yard = _yardstick(ref, precision, lambda m: m.criterion(_ac(lambda: m.ptn_shared(x)).float(), y))
    worst = grads_close(mod, ref, TOL[precision], "ptn_shared ", skip=("mlp_encoder", "encoder_layers"), yard=yard)
    print("worst grad", precision, worst)
    with pytest.raises(ValueError):                          # E + 1 tokens must fit PositionalEncoding's max_len
        short = api.SimpleTransformer(precision=precision, **dict(cfg, seq_len=2)).to(DEV)
        short.ptn_shared(torch.randn(B, 2, 3, 256, device=DEV))
    with pytest.raises(ValueError):                          # more clips than CLS batch slots
        mod.ptn(torch.randn(B + 1, 16, 3, 256, device=DEV))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shared_parameters_complete_only_after_their_last_contribution(api, precision):
    """A module applied several times per step (ptn_shared: transformer_encoder0 once per expert) writes the same
    gradient views several times through the direct sinks: the reducer must arm a bucket's all-reduce only after the
    LAST contribution (ddp.note_use / _on_direct), and the accumulated gradients must equal autograd's."""
    from tvt_b200 import ddp
    cfg = dict(batch_size=8, seq_len=12, cls=1, dropout=0.0, input_dimension=64, nhead=2, nhid=128, nlayers=2,
               model="ptn_shared", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15, precision=precision)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(8, 12, 3, 64, generator=gen).to(DEV)
    y = _targets(8, 15, gen).to(DEV)
    grads = {}
    for direct in (False, True):
        torch.manual_seed(1130)
        mod = api.SimpleTransformer(**cfg).to(DEV).train()
        params = [p for n, p in mod.named_parameters() if not n.startswith(("mlp_encoder", "encoder_layers"))]
        red = ddp.GradBucketReducer(params, bucket_bytes=1 << 15, direct=direct)
        fired = []
        if direct:
            orig = red._on_grad
            enc0 = {id(p) for p in mod.transformer_encoder0.parameters()}

            def spy(p, orig=orig):
                if id(p) in enc0:
                    fired.append(red._uses.get(id(p), 0))
                return orig(p)
            red._on_grad = spy
        red.zero_grad()
        loss = mod._loss(mod.ptn_shared(x), y)
        if direct:      # forward registered E = 3 uses of every shared-encoder parameter
            w0 = mod.transformer_encoder0.layers[0].linear1.weight
            assert red._uses[id(w0)] == 3
        loss.backward()
        if direct:
            assert fired and all(left == 0 for left in fired), "a shared parameter completed before its last contribution"
            assert all(b["pending"] == 0 for b in red.buckets)
        red.finish()
        grads[direct] = {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}
        red.remove()
    for n in grads[False]:
        assert_close(grads[True][n], grads[False][n], 1e-5, "shared direct " + n)      # fp32 atomics: summation order varies


def test_graphed_folded_inference_recaptures_after_a_weight_update(api):
    """Same as below at a size whose evaluation forward takes the LayerNorm-free stack: the graph reads the CACHED folded
    weights (W diag(gamma), c, b'), which a weight OR LayerNorm-parameter update invalidates."""
    from tvt_b200 import ops
    kw = dict(in_dims=(256,), d=512, nhead=8, nhid=2048, nlayers=2, dropout=0.0, batch_size=256, frames=32, n_classes=15, fusion="sum")
    torch.manual_seed(1130)
    mod = api.FusionTransformer(precision="bf16", **kw).to(DEV).eval()
    assert ops.ln_fold_supported(256 * 33, 512, 512)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(256, 32, 256, generator=gen).to(DEV)
    graphed = api.GraphedForward(lambda t: mod([t])[0], [x], module=mod)
    before = graphed(x).clone()
    assert graphed.captures == 1
    with torch.no_grad():
        mod.streams[0].transformer_encoder.layers[0].norm1.weight.mul_(1.5)      # only a LayerNorm gamma changes
        eager = mod([x])[0].clone()
    got = graphed(x)
    torch.cuda.synchronize()
    assert graphed.captures == 2
    assert torch.equal(got, eager) and not torch.equal(got, before)


def test_graphed_forward_recaptures_after_a_weight_update(api):
    """the captured graph holds raw addresses of the cached bf16 weight planes; a plain torch optimizer step
    (what the reference's configure_optimizers returns) re-splits the weights into NEW planes, so the replay must notice
    and re-capture instead of silently using stale / freed memory."""
    cfg = dict(batch_size=8, seq_len=16, cls=1, dropout=0.0, input_dimension=256, nhead=4, nhid=512, nlayers=2,
               model="ptn", learning_rate=1e-1, momentum=0.0, weight_decay=0.0, n_classes=15)
    torch.manual_seed(1130)
    mod = api.SimpleTransformer(precision="bf16", **cfg).to(DEV)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, 3, 256, generator=gen).to(DEV)
    y = _targets(8, 15, gen).to(DEV)
    mod.eval()
    graphed = api.GraphedForward(mod.ptn, [x])
    before = graphed(x).clone()
    assert graphed.captures == 1
    graphed(x)
    assert graphed.captures == 1                              # nothing changed: no re-capture
    opt = mod.configure_optimizers()                          # torch.optim.SGD, as the reference returns
    mod.train()
    mod.training_step({"experts": x, "label": y}, 0).backward()
    opt.step()
    mod.eval()
    with torch.no_grad():
        eager = mod.ptn(x).clone()
    got = graphed(x)
    torch.cuda.synchronize()
    assert graphed.captures == 2
    assert torch.equal(got, eager) and not torch.equal(got, before)
    mod.load_state_dict({k: v.clone() for k, v in mod.state_dict().items()})   # in-place copy: versions bump again
    got = graphed(x)
    torch.cuda.synchronize()
    assert graphed.captures == 3 and torch.equal(got, eager)


def test_drop_in_at_reference_width_matches_golden(api):
    """Same constructor call as the reference (d = 2048 hard-coded there) reproduces the frozen reference logits."""
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.pt"), weights_only=False)
    g = gold["ptn"]
    torch.manual_seed(gold["seed"])
    mod = api.SimpleTransformer(precision="fp32", **g["cfg"]).to(DEV)        # same RNG stream as the reference
    assert sorted(mod.state_dict().keys()) == g["state_dict_keys"]
    gen = torch.Generator().manual_seed(gold["seed"])
    x = torch.randn(2, 4, 3, 2048, generator=gen).to(DEV)
    y = (torch.rand(2, 15, generator=gen) < 0.15).float().to(DEV)
    logits = mod.ptn(x)
    assert_close(logits, g["logits"], 1e-3, "golden logits")
    loss = mod.training_step({"experts": x, "label": y}, 0)
    assert_close(loss, g["loss"], 1e-3, "golden loss")
    loss.backward()
    for name, p in mod.named_parameters():
        if name in g["grad_norms"] and float(g["grad_norms"][name]) > 0:
            got = p.grad.double().norm().item()
            want = float(g["grad_norms"][name])
            assert abs(got - want) <= 2e-3 * want, (name, got, want)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("fusion,pyramid", [("sum", False), ("cross", False), ("cross", True)])
def test_fusion_transformer_parity(api, precision, fusion, pyramid):
    from oracle import param
    B = 6 if precision == "fp32" else 128
    kw = dict(in_dims=(2048, 1024, 128), d=256, nhead=4, nhid=512, nlayers=2, dropout=0.0, batch_size=B, frames=20,
              n_classes=15, fusion=fusion, pyramid=pyramid)
    torch.manual_seed(1130)
    ref = param.FusionTransformer(**kw).to(DEV)
    mod = copy_state(api.FusionTransformer(precision=precision, **kw), ref).to(DEV)
    _no_dropout(ref, mod)
    gen = torch.Generator().manual_seed(1130)
    xs = [torch.relu(torch.randn(B, 20, D, generator=gen) * 0.5).to(DEV) if D > 128 else torch.randn(B, 20, D, generator=gen).to(DEV)
          for D in kw["in_dims"]]
    y = _targets(B, 15, gen).to(DEV)
    logits_r, pyr_r = ref(xs)
    loss_r = torch.nn.functional.binary_cross_entropy_with_logits(logits_r, y)
    if pyramid:
        loss_r = loss_r + torch.nn.functional.binary_cross_entropy(pyr_r, y)
    loss_r.backward()
    logits, prob, ploss = mod(xs, y)
    from tvt_b200.functions import DistillLossFn
    loss = DistillLossFn.apply(logits, None, y, 1.0, 0.0, 0.0, 1.0)[0]
    if pyramid:
        loss = loss + ploss[0]
        assert_close(prob, pyr_r, TOL[precision], "pyramid probs")
    loss.backward()
    assert_close(logits, logits_r, TOL[precision], "logits")
    assert_close(loss, loss_r, TOL[precision], "loss")
    def run(m):
        lg, pr = _ac(lambda: m(xs))
        l = torch.nn.functional.binary_cross_entropy_with_logits(lg.float(), y)
        return l + torch.nn.functional.binary_cross_entropy(pr.float().clamp(1e-6, 1 - 1e-6), y) if pyramid else l

    worst = grads_close(mod, ref, TOL[precision], f"{fusion} ", yard=_yardstick(ref, precision, run))
    print("worst grad", precision, fusion, pyramid, worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_distillation_step_parity(api, precision):
    """Frozen 3-expert cross-attention teacher -> RGB-only pyramid student, BCE + CE + KL + pyramid BCE."""
    from oracle import param
    B = 8 if precision == "fp32" else 128
    common = dict(d=256, nhead=4, nhid=512, nlayers=2, dropout=0.0, batch_size=B, frames=16, n_classes=15)
    torch.manual_seed(1130)
    t_ref = param.FusionTransformer(in_dims=(2048, 1024, 128), fusion="cross", **common).to(DEV).eval()
    s_ref = param.FusionTransformer(in_dims=(2048,), fusion="sum", pyramid=True, **common).to(DEV)
    teacher = copy_state(api.FusionTransformer(in_dims=(2048, 1024, 128), fusion="cross", precision=precision, **common), t_ref).to(DEV)
    student = copy_state(api.FusionTransformer(in_dims=(2048,), fusion="sum", pyramid=True, precision=precision, **common), s_ref).to(DEV)
    with torch.no_grad():          # decisive teacher: hard labels (argmax) must not hinge on bf16 rounding of near-ties
        for t in (t_ref, teacher):
            t.mlp_head[1].bias[3] += 2.0
    _no_dropout(t_ref, s_ref, teacher, student)
    trainer = api.DistillationTrainer(teacher, student, temperature=2.0, alpha=0.5).train()
    gen = torch.Generator().manual_seed(1130)
    xs = [torch.randn(B, 16, D, generator=gen).to(DEV) for D in (2048, 1024, 128)]
    y = _targets(B, 15, gen).to(DEV)
    with torch.no_grad():
        t_logits, _ = t_ref(xs)
    s_logits, s_pyr = s_ref(xs[:1])
    loss_r, _ = param.distill_loss(s_logits, t_logits, y, temperature=2.0, alpha=0.5, pyramid=s_pyr)
    loss_r.backward()
    loss = trainer.training_step({"experts": xs, "label": y})
    loss.backward()
    assert_close(loss, loss_r, TOL[precision], "distillation loss")
    def run(m):
        lg, pr = _ac(lambda: m(xs[:1]))
        return param.distill_loss(lg.float(), t_logits, y, temperature=2.0, alpha=0.5, pyramid=pr.float().clamp(1e-6, 1 - 1e-6))[0]

    grads_close(student, s_ref, TOL[precision], "student ", yard=_yardstick(s_ref, precision, run))
    assert all(p.grad is None for p in teacher.parameters())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frame_stream_parity(api, precision):
    """FrameTransformer's real widths: d = 896, 2 heads (head_dim 448), ff = 512, 14 scene tokens, GELU head."""
    from oracle import param
    kw = dict(d=896, nhead=2, nhid=512, nlayers=2, dropout=0.0, seq_len=14, n_classes=19)
    torch.manual_seed(1130)
    ref = param.FrameStream(**kw).to(DEV)
    mod = copy_state(api.FrameStream(precision=precision, **kw), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    B = 4 if precision == "fp32" else 128
    feats = torch.randn(B, 14, 896, generator=gen).to(DEV)
    y = _targets(B, 19, gen).to(DEV)
    teacher = torch.randn(B, 19, generator=gen).to(DEV)
    loss_r, _ = param.distill_loss(ref(feats), teacher, y)
    loss_r.backward()
    loss = mod.training_step((y, feats), 0, teacher_logits=teacher)
    loss.backward()
    assert_close(mod(feats), ref(feats), TOL[precision], "logits")
    assert_close(loss, loss_r, TOL[precision], "loss")
    grads_close(mod, ref, TOL[precision], "frame ", yard=_yardstick(ref, precision, lambda m: param.distill_loss(_ac(lambda: m(feats)).float(), teacher, y)[0]))
    # seq-first reference signature of TransformerBase
    x = torch.randn(14, 4, 896, generator=gen).to(DEV)
    assert_close(mod.distil_transformer(x), ref.distil_transformer(x), TOL[precision], "TransformerBase.forward")


@pytest.mark.parametrize("precision,b,n", [("fp32", 4, 50), ("bf16", 64, 50), ("bf16", 8, 197)])
def test_vit_prenorm_transformer_parity(api, precision, b, n):
    """src/models/vit.py Transformer at ViViT's defaults (dim 192, 3 heads x 64, GELU MLP 4x, pre-norm)."""
    from oracle import param
    torch.manual_seed(1130)
    ref = param.VitTransformer(192, 2, 3, 64, 768).to(DEV)
    mod = copy_state(api.vit.Transformer(192, 2, 3, 64, 768, precision=precision), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    x = torch.randn(b, n, 192, generator=gen).to(DEV)
    w = torch.randn(b, n, 192, generator=gen).to(DEV)
    (ref(x) * w).sum().backward()
    out = mod(x)
    (out.float() * w).sum().backward()
    assert_close(out, ref(x), TOL[precision], "vit out")
    grads_close(mod, ref, TOL[precision], "vit ", yard=_yardstick(ref, precision, lambda m: (_ac(lambda: m(x)).float() * w).sum()))


def test_vit_matches_reference_golden(api):
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.pt"), weights_only=False)
    torch.manual_seed(gold["seed"])
    mod = api.vit.Transformer(32, 2, 2, 16, 64, precision="fp32").to(DEV).eval()   # same RNG stream as the reference
    gen = torch.Generator().manual_seed(gold["seed"])
    x = torch.randn(2, 5, 32, generator=gen).to(DEV)
    assert_close(mod(x), gold["vit"]["out"], 1e-3, "vit golden")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_token_injection_fusion_parity(api, precision):
    """FrameTransformer "sum" mode: the other modality's CLS vector joins the sequence as an extra token."""
    from oracle import param
    B = 4 if precision == "fp32" else 96
    kw = dict(d=256, nhead=4, nhid=512, nlayers=2, dropout=0.0, seq_len=15, n_classes=19)
    torch.manual_seed(1130)
    ref = param.FrameStream(**kw).to(DEV)
    mod = copy_state(api.FrameStream(precision=precision, **kw), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    feats = torch.randn(B, 14, 256, generator=gen).to(DEV)
    other = torch.randn(B, 256, generator=gen).to(DEV)
    y = _targets(B, 19, gen).to(DEV)
    lr = torch.nn.functional.binary_cross_entropy_with_logits(ref.sum_forward(feats, other), y)
    lr.backward()
    logits = mod.sum_forward(feats, other)
    from tvt_b200.functions import DistillLossFn
    loss = DistillLossFn.apply(logits, None, y, 1.0, 0.0, 0.0, 1.0)[0]
    loss.backward()
    assert_close(logits, ref.sum_forward(feats, other), TOL[precision], "sum logits")
    assert_close(loss, lr, TOL[precision], "loss")
    grads_close(mod, ref, TOL[precision], "sum ", yard=_yardstick(
        ref, precision, lambda m: torch.nn.functional.binary_cross_entropy_with_logits(_ac(lambda: m.sum_forward(feats, other)).float(), y)))


def test_reasoning_and_spatial_pyramid_modules(api):
    from oracle import param
    torch.manual_seed(1130)
    ref = param.Reasoning(num_segments=4, num_frames=5, num_class=15, img_dim=896).to(DEV).eval()
    mod = copy_state(api.Reasoning(num_segments=4, num_frames=5, num_class=15, img_dim=896, precision="fp32"), ref).to(DEV).eval()
    gen = torch.Generator().manual_seed(1130)
    x = torch.randn(3, 20, 896, generator=gen).to(DEV)
    assert_close(mod(x), ref(x), 1e-3, "Reasoning")
    assert_close(api.sum_group(x, 3), param.sum_group(x, 3), 1e-6, "sum_group")
    sp_ref = param.SpatialPyramid().to(DEV)
    sp = copy_state(api.SpatialPyramid(), sp_ref).to(DEV)
    maps = [torch.randn(4, c, s, s, generator=gen).to(DEV) for c, s in ((128, 28), (256, 14), (512, 7))]
    assert_close(sp(*maps), sp_ref(*maps), 1e-3, "SpatialPyramid")


def test_training_mode_dropout_runs_and_is_consistent(api):
    """dropout > 0: forward/backward run, are finite, and the recomputed masks make the step self-consistent:
    a directional finite difference of the loss (same seeds) matches the analytic gradient."""
    from tvt_b200 import functions
    kw = dict(in_dims=(256,), d=256, nhead=4, nhid=512, nlayers=1, dropout=0.5, batch_size=8, frames=16, n_classes=15, fusion="sum")
    torch.manual_seed(3)
    mod = api.FusionTransformer(precision="fp32", **kw).to(DEV).train()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(8, 16, 256, generator=gen).to(DEV)
    y = _targets(8, 15, gen).to(DEV)

    def loss_at():
        functions._seed_counter[0] = 1000          # replay the same dropout masks
        logits, _, _ = mod([x], None)
        return functions.DistillLossFn.apply(logits, None, y, 1.0, 0.0, 0.0, 1.0)[0]

    l0 = loss_at()
    l0.backward()
    assert torch.isfinite(l0)
    w = mod.streams[0].transformer_encoder.layers[0].linear2.weight
    gdir = torch.randn(w.shape, generator=gen).to(DEV)
    gdir /= gdir.norm()
    analytic = float((w.grad * gdir).sum())
    eps = 1e-2
    with torch.no_grad():
        w.add_(eps * gdir); lp = float(loss_at()); w.sub_(2 * eps * gdir); lm = float(loss_at()); w.add_(eps * gdir)
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - analytic) <= 0.05 * max(abs(analytic), 1e-4) + 1e-5, (fd, analytic)
    # eval mode is deterministic and differs from train mode
    mod.eval()
    a, b = mod([x])[0], mod([x])[0]
    # fp32 mode: the forward GEMMs sum short tensor-core chains with fp32 atomics (ops.Mode.linear_fwd), whose order varies
    # from launch to launch: equal to a few fp32 ulp, not bit for bit (the bf16 mode is bit-deterministic, see below)
    assert_close(a, b, 2e-6, "eval determinism (fp32 mode)")
    mod16 = api.FusionTransformer(precision="bf16", **kw).to(DEV).eval()
    assert torch.equal(mod16([x])[0], mod16([x])[0])


@pytest.mark.parametrize("precision", ["bf16"])
def test_c5_shapes_distillation_parity(api, precision):
    """The flagship shapes of BASELINE config 5 (128 frames -> S = 129 with a CUDA-core tail row / tail key,
    d = 768, 12 heads of 64, ff = 3072, pyramid groups 2/3/4 over 128 frames, cross-attention over 258 keys)
    at reduced depth (2 layers) and batch (32 clips): loss and every student gradient against the fp32 oracle."""
    from oracle import param
    B = 32
    common = dict(d=768, nhead=12, nhid=3072, nlayers=2, dropout=0.0, batch_size=B, frames=128, n_classes=15)
    torch.manual_seed(1130)
    t_ref = param.FusionTransformer(in_dims=(2048, 1024, 128), fusion="cross", **common).to(DEV).eval()
    s_ref = param.FusionTransformer(in_dims=(2048,), fusion="sum", pyramid=True, **common).to(DEV)
    teacher = copy_state(api.FusionTransformer(in_dims=(2048, 1024, 128), fusion="cross", precision=precision, **common), t_ref).to(DEV)
    student = copy_state(api.FusionTransformer(in_dims=(2048,), fusion="sum", pyramid=True, precision=precision, **common), s_ref).to(DEV)
    with torch.no_grad():
        for t in (t_ref, teacher):
            t.mlp_head[1].bias[3] += 2.0
    _no_dropout(t_ref, s_ref, teacher, student)
    trainer = api.DistillationTrainer(teacher, student, temperature=2.0, alpha=1.0).train()
    gen = torch.Generator().manual_seed(1130)
    xs = [torch.relu(torch.randn(B, 128, D, generator=gen) * 0.5).to(DEV) if D > 128 else torch.randn(B, 128, D, generator=gen).to(DEV)
          for D in (2048, 1024, 128)]
    y = _targets(B, 15, gen).to(DEV)
    with torch.no_grad():
        t_logits, _ = t_ref(xs)
        mine_t, _, _ = teacher(xs)
    assert_close(mine_t, t_logits, TOL[precision], "teacher logits (cross-attention over 258 keys)")
    s_logits, s_pyr = s_ref(xs[:1])
    loss_r, _ = param.distill_loss(s_logits, t_logits, y, temperature=2.0, alpha=1.0, pyramid=s_pyr)
    loss_r.backward()
    loss = trainer.training_step({"experts": xs, "label": y})
    loss.backward()
    assert_close(loss, loss_r, TOL[precision], "C5-shape distillation loss")

    def run(m):
        lg, pr = _ac(lambda: m(xs[:1]))
        return param.distill_loss(lg.float(), t_logits, y, temperature=2.0, alpha=1.0, pyramid=pr.float().clamp(1e-6, 1 - 1e-6))[0]

    worst = grads_close(student, s_ref, TOL[precision], "C5 student ", yard=_yardstick(s_ref, precision, run))
    print("C5-shape worst grad", worst)


def test_full_size_properties(api):
    """Size-independent properties at BASELINE config 5's FULL per-GPU size (256 clips x 128 frames, d = 768,
    12 layers, 12 heads): clips are independent (permuting the batch permutes the logits), evaluation is
    deterministic, and train mode with every dropout at 0 equals eval mode."""
    kw = dict(in_dims=(2048,), d=768, nhead=12, nhid=3072, nlayers=12, dropout=0.0, batch_size=256, frames=128, n_classes=15,
              fusion="sum", pyramid=True)
    torch.manual_seed(1130)
    mod = api.FusionTransformer(precision="bf16", **kw).to(DEV).eval()
    with torch.no_grad():                      # identical CLS in every batch slot so that clips are exchangeable
        mod.streams[0].cls.copy_(mod.streams[0].cls[:, :1].expand_as(mod.streams[0].cls))
    gen = torch.Generator().manual_seed(1130)
    x = torch.relu(torch.randn(256, 128, 2048, generator=gen) * 0.5).to(DEV)
    perm = torch.randperm(256, generator=gen).to(DEV)
    with torch.no_grad():
        a, pa, _ = mod([x])
        b, pb, _ = mod([x])
        c, pc, _ = mod([x[perm]])
        _no_dropout(mod)
        mod.train()
        d_, pd_, _ = mod([x])
    # the encoder path is bit-deterministic; the pyramid's first Linear runs split-K with fp32 atomics, whose
    # summation order varies from launch to launch: a few fp32 ulp, which the bf16 rounding of the hidden
    # layer can turn into one bf16 ulp (0.4 %) of a hidden unit
    close = lambda u, v: torch.allclose(u, v, rtol=0, atol=2e-3)
    assert torch.equal(a, b) and close(pa, pb)                             # deterministic
    assert torch.equal(c, a[perm]) and close(pc, pa[perm])                 # clip independence
    # dropout-free train == eval up to bf16 round-off: evaluation under no_grad takes the LayerNorm-free stack
    # (hostapi.common.run_encoder_folded), train mode the layer-by-layer path - two arithmetic routes to the same numbers
    assert_close(d_, a, 2e-2, "train (no dropout) vs eval logits")
    assert torch.allclose(pd_, pa, rtol=0, atol=1e-2)
    assert torch.isfinite(a).all() and float(pa.min()) >= 0.0 and float(pa.max()) <= 1.0


def test_eval_buffer_matches_reference_eval_path(api):
    """Evaluation read-out (transformer.py:146-158, callbacks.py:34-45): sigmoid / label cast / per-threshold
    predictions / top-1 from ONE launch per batch into running device buffers, against the oracle restatement.
    Integer outputs are exact; a thresholded bit may only differ where the probability sits on the threshold."""
    from oracle import param
    gen = torch.Generator().manual_seed(1130)
    C = 19
    buf = api.EvalBuffer(capacity=700, n_classes=C)
    all_logits, all_tgt = [], []
    for B in (256, 1, 300, 37):                                # ragged batches, the last one partial
        logits = torch.randn(B, C, generator=gen) * 3.0
        logits[0, 3] = logits[0, 7] = logits[0].max() + 1.0     # a tie: torch.argmax takes the first
        tgt = _targets(B, C, gen).double().reshape(B, 1, C)     # the loader's [B, 1, C] float64 labels
        buf.append(logits.to(DEV), tgt.to(DEV))
        all_logits.append(logits)
        all_tgt.append(tgt)
    logits, tgt = torch.cat(all_logits), torch.cat(all_tgt)
    probs, labels, preds, top1 = param.eval_readout(logits, tgt)
    assert buf.rows == 594 and buf.probs.shape == (594, C)
    assert_close(buf.probs.cpu(), probs, 1e-6, "probs")
    assert torch.equal(buf.labels.cpu(), labels)
    assert torch.equal(buf.top1.cpu().long(), top1)
    for t, ref in zip(api.REFERENCE_THRESHOLDS, preds):
        got = buf.predictions(t).cpu()
        off = (got != ref) & ((probs - t).abs() > 1e-6)
        assert not off.any(), f"threshold {t}: {int(off.sum())} predictions differ away from the threshold"
    host = buf.to_host()
    assert host["probs"].shape == (594, C) and host["labels"].dtype.name == "int32"
    with pytest.raises(ValueError):
        buf.append(torch.zeros(200, C, device=DEV))             # would overflow the capacity
    buf.reset()
    assert buf.rows == 0 and buf.probs.shape[0] == 0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_inference_forward_is_bit_identical(api, precision):
    """The CUDA-graph replay of the inference-only forward returns exactly what the eager forward returns, for
    fresh inputs after capture, and matches the oracle within the precision's tolerance."""
    from oracle import param
    B = 8
    cfg = dict(batch_size=B, seq_len=16, cls=1, dropout=0.5, input_dimension=256, nhead=4, nhid=512, nlayers=2,
               model="ptn", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15)
    torch.manual_seed(1130)
    ref = param.SimpleTransformer(**cfg).to(DEV).eval()
    mod = copy_state(api.SimpleTransformer(precision=precision, **cfg), ref).to(DEV).eval()
    gen = torch.Generator().manual_seed(7)
    x0 = torch.randn(B, 16, 3, 256, generator=gen).to(DEV)
    graphed = api.GraphedForward(mod.ptn, [x0])
    for _ in range(3):
        x = torch.randn(B, 16, 3, 256, generator=gen).to(DEV)
        with torch.no_grad():
            eager = mod.ptn(x).clone()
            want = ref.ptn(x)
        got = graphed(x)
        torch.cuda.synchronize()
        if precision == "bf16":
            assert torch.equal(got, eager)
        else:                      # fp32 mode sums split-K partials with fp32 atomics: equal to a few ulp
            assert_close(got, eager, 2e-6, "graph vs eager (fp32 mode)")
        assert_close(got, want, TOL[precision], "graphed logits")
    with pytest.raises(ValueError):
        graphed(torch.zeros(B + 1, 16, 3, 256, device=DEV))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_collaborative_gating_parity(api, precision):
    """Collaborative gating (collabgating.py:3-56) as three stacked tensor-core GEMMs against the oracle (which is
    pinned to the unmodified reference class by tests/golden): forward on the stacked and on the reference's nested
    list layout, and every parameter gradient."""
    from oracle import param
    torch.manual_seed(1130)
    ref = param.CollaborativeGating().to(DEV)
    mod = copy_state(api.CollaborativeGating(precision=precision), ref).to(DEV)
    gen = torch.Generator().manual_seed(3)
    B, S = (4, 3) if precision == "fp32" else (16, 8)
    xs = [torch.randn(B, S, D, generator=gen).to(DEV) for D in (2048, 1024, 128)]
    w = torch.randn(B, S, 1024, generator=gen).to(DEV)
    out_ref = ref(xs)
    (out_ref * w).sum().backward()
    out = mod(xs)
    (out * w).sum().backward()
    assert_close(out, out_ref, TOL[precision], "collab out")
    nested = [[[x[b, s].reshape(1, -1) for x in xs] for s in range(S)] for b in range(B)]
    with torch.no_grad():
        if precision == "bf16":
            assert torch.equal(mod(nested), mod(xs))
        else:
            assert_close(mod(nested), mod(xs), 2e-6, "nested vs stacked input (fp32 mode: atomic summation order)")
    yard = _yardstick(ref, precision, lambda m: (_ac(lambda: m(xs)).float() * w).sum())
    grads_close(mod, ref, TOL[precision], "collab ", yard=yard)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_direct_gradient_sinks_match_autograd_accumulation(api, precision):
    """With a GradBucketReducer(direct=True) the backward kernels accumulate straight into the flat gradient buckets
    (no temporaries, zero-fills or add_ launches): same gradients as autograd's accumulation, every bucket's
    all-reduce trigger still fires, and a second backward without zero_grad accumulates (torch semantics)."""
    from tvt_b200 import ddp
    kw = dict(in_dims=(64, 32), d=64, nhead=2, nhid=128, nlayers=2, dropout=0.0, batch_size=8, frames=12, n_classes=15,
              fusion="cross", pyramid=True, precision=precision)
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(8, 12, D, generator=gen).to(DEV) for D in (64, 32)]
    y = _targets(8, 15, gen).to(DEV)
    grads = {}
    for direct in (False, True):
        torch.manual_seed(1130)
        mod = _no_dropout(api.FusionTransformer(**kw)).to(DEV).train()
        red = ddp.GradBucketReducer(list(mod.parameters()), bucket_bytes=1 << 16, direct=direct)
        assert len(red.buckets) >= 2
        red.zero_grad()
        logits, _, ploss = mod(xs, y)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y) + ploss[0]
        loss.backward()
        used = [b for b in red.buckets if any(id(p) in b["seen"] for p in b["params"])]
        assert all(b["pending"] == len(b["params"]) - len(b["seen"]) for b in red.buckets)
        assert sum(len(b["seen"]) for b in red.buckets) >= 0.8 * len(red.params) and used
        red.finish()
        grads[direct] = {n: p.grad.clone() for n, p in mod.named_parameters()}
        if direct:
            logits, _, ploss = mod(xs, y)                      # second backward, no zero_grad: accumulates
            (torch.nn.functional.binary_cross_entropy_with_logits(logits, y) + ploss[0]).backward()
            for n, p in mod.named_parameters():
                assert_close(p.grad, 2 * grads[True][n], 1e-5 if precision == "fp32" else 1e-3, "accumulated " + n)
        red.remove()
    for n in grads[False]:
        assert_close(grads[True][n], grads[False][n], 1e-5 if precision == "fp32" else 1e-6, "direct " + n)
    assert ddp.direct_target(next(mod.parameters())) is None   # removed reducers no longer capture gradients


@pytest.mark.parametrize("pyramid", [False, True])
def test_graphed_training_step_matches_eager_steps(api, pyramid):
    """Whole-step CUDA graph (zero_grad + teacher forward + student forward / backward + flat-bucket AdamW) replayed twice
    == the same two steps launched eagerly from the same state: dropout p = 0.5 everywhere, so this also proves that the
    masks of a replay come from the DEVICE step counter (a graph that replayed baked seeds would repeat step 1's masks in
    step 2) and that forward and backward of a replay regenerate the same masks.  Parameters / losses agree to fp32
    round-off of the split-K atomics, not bit for bit."""
    from tvt_b200 import ddp, optim, functions
    B = 16
    common = dict(d=128, nhead=2, nhid=256, nlayers=2, dropout=0.5, batch_size=B, frames=16, n_classes=15, precision="bf16")
    torch.manual_seed(1130)
    teacher = api.FusionTransformer(in_dims=(256, 64), fusion="cross", **common).to(DEV)
    student = api.FusionTransformer(in_dims=(256,), fusion="sum", pyramid=pyramid, **common).to(DEV)
    trainer = api.DistillationTrainer(teacher, student, temperature=2.0, alpha=1.0).train()
    red = ddp.GradBucketReducer([p for p in student.parameters()], bucket_bytes=1 << 18, average=False)
    opt = optim.FlatOptimizer(red, modes=[student.mode], kind="adamw", lr=1e-3, weight_decay=0.01)
    gen = torch.Generator().manual_seed(3)
    batches = [([torch.randn(B, 16, D, generator=gen).to(DEV).bfloat16() for D in (256, 64)], _targets(B, 15, gen).to(DEV)) for _ in range(2)]

    def step(xs, y):
        red.zero_grad()
        loss = trainer.training_step({"experts": xs, "label": y})
        loss.backward()
        red.finish()
        opt.step()
        return loss

    stepper = api.GraphedTrainStep(step, opt, batches[0])
    seeds = [stepper.prepare(*batches[0]), stepper.prepare(*batches[1])]     # host seed counter baked into each graph
    state = [(b["p"].clone(), b["m"].clone(), b["v"].clone()) for b in opt.buckets]
    ctr0 = stepper.counters.clone()
    graph_losses = [float(stepper(*batches[i])) for i in range(2)]
    graph_params = [b["p"].clone() for b in opt.buckets]
    assert stepper.kernels_per_replay > 50 and len(stepper.graphs) == 2
    assert int(stepper.counters[0]) == int(ctr0[0]) + 2
    # restore the state and run the same two steps eagerly (same baked host seeds, same device counter values)
    for b, (p_, m_, v_) in zip(opt.buckets, state):
        b["p"].copy_(p_); b["m"].copy_(m_); b["v"].copy_(v_)
        if b["hi"] is not None:
            b["hi"].copy_(p_.bfloat16())
    stepper.counters.copy_(ctr0)
    eager_losses = []
    for i in range(2):
        functions._seed_counter[0] = seeds[i]
        stepper._advance()
        eager_losses.append(float(step(*batches[i])))
    for a, b in zip(graph_losses, eager_losses):
        assert abs(a - b) <= 1e-3 * abs(b), (graph_losses, eager_losses)   # the order of the wgrad atomics differs run to run (more so with dynamic tile claims)
    assert abs(graph_losses[0] - graph_losses[1]) > 1e-3 * abs(graph_losses[0])          # different batches / masks
    for gp, b in zip(graph_params, opt.buckets):
        assert_close(gp, b["p"], 1e-5, "parameters after two replays vs two eager steps")
    # a replay on the SAME batch draws new masks each time (the device counter advanced)
    l1, l2 = float(stepper(*batches[0])), float(stepper(*batches[0]))
    assert l1 != l2
    stepper.close()
    red.remove()


@pytest.mark.parametrize("d,H,ff,L,B,T", [(512, 8, 2048, 3, 256, 32), (768, 12, 3072, 2, 256, 128)])
def test_layernorm_free_inference_stack(api, d, H, ff, L, B, T):
    """Inference-only encoder stacks run WITHOUT LayerNorm kernels (hostapi.common.run_encoder_folded: statistics from the
    producing GEMM's epilogue, gamma folded into the consuming weights, LN-as-residual recomputed in the epilogue): same
    result as the layer-by-layer path within bf16 round-off, and both within the bf16 bar of the fp32 torch encoder."""
    from tvt_b200 import ops
    from tvt_b200.hostapi import common
    torch.manual_seed(1130)
    enc = common.make_encoder(d, H, ff, 0.0, L).to(DEV).eval()
    with torch.no_grad():                      # non-trivial LayerNorm affine parameters and distinct layers
        for i, layer in enumerate(enc.layers):
            for nrm in (layer.norm1, layer.norm2):
                nrm.weight.add_(0.2 * torch.randn_like(nrm.weight))
                nrm.bias.add_(0.1 * torch.randn_like(nrm.bias))
            layer.linear1.weight.mul_(1.0 + 0.05 * i)
    S = T + 1
    n = B * S
    assert all(ops.ln_fold_supported(n, N_, K_) for N_, K_ in ((3 * d, d), (d, d), (ff, d), (d, ff)))
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(S, B, d, generator=gen).to(DEV)
    tok = x.transpose(0, 1).reshape(n, d).to(torch.bfloat16).contiguous()
    mode = ops.Mode("bf16")
    import tvt_b200
    with torch.no_grad():
        want = enc(x).transpose(0, 1).reshape(n, d)
        for flag in (True, False):             # first use converts the weights (cached afterwards): keep it out of the launch counts
            common.FOLD_LAYERNORM = flag
            common.run_encoder(mode, enc, tok, B, False)
        l0 = tvt_b200.capi.launches
        common.FOLD_LAYERNORM = True
        folded = common.run_encoder(mode, enc, tok, B, False)
        l1 = tvt_b200.capi.launches
        common.FOLD_LAYERNORM = False
        plain = common.run_encoder(mode, enc, tok, B, False)
        l2 = tvt_b200.capi.launches
        common.FOLD_LAYERNORM = True
    assert (l1 - l0) == 5 * L + 1 and (l2 - l1) == 7 * L            # 2 L LayerNorm launches became 1
    assert_close(folded, plain, 1e-2, "folded vs layer-by-layer")
    assert_close(plain, want, 2e-2, "layer-by-layer vs torch fp32")
    assert_close(folded, want, 2e-2, "folded vs torch fp32")
    # with autograd recording, the stack must take the differentiable path
    out = common.run_encoder(mode, enc, tok.clone().requires_grad_(True), B, False)
    assert out.requires_grad


def test_cross_memory_built_without_concatenation(api):
    """Inference through a 3-expert cross-attention FusionTransformer (src/models/transformer.py:110-121): the memory experts' last
    LayerNorm writes its tokens straight into its block of the per-clip memory (tvt_layernorm_fwd blocked output).  Same bits as
    laying the dense tokens into the buffer by a copy, same launches minus the copies, and the bf16 bar against the differentiable
    (torch.cat) path."""
    import tvt_b200
    from tvt_b200 import hostapi, ops
    from tvt_b200.hostapi import fusion
    torch.manual_seed(7)
    B, T, d = 256, 32, 512
    model = hostapi.FusionTransformer((2048, 1024, 128), d=d, nhead=8, nhid=2048, nlayers=2, dropout=0.0, batch_size=B, frames=T,
                                      n_classes=15, fusion="cross", precision="bf16").to(DEV).eval()
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(B, T, D, generator=gen).to(DEV) for D in (2048, 1024, 128)]
    with torch.no_grad():
        model(xs)                                                # weight conversions out of the launch counts
        l0 = tvt_b200.capi.launches
        blocked = model(xs)[0]
        l1 = tvt_b200.capi.launches
        ok = fusion.blocked_output_ok
        fusion.blocked_output_ok = lambda *a, **k: False         # same forward, tokens copied into the buffer afterwards
        try:
            copied = model(xs)[0]
        finally:
            fusion.blocked_output_ok = ok
        l2 = tvt_b200.capi.launches
    assert torch.equal(blocked, copied), "blocked LayerNorm output differs from dense output + copy"
    assert l1 - l0 == l2 - l1                                   # same tvt launches; the copy path adds two ATen copies on top
    with torch.enable_grad():
        ref = model(xs)[0]                                       # differentiable path: EncoderLayerFn per layer + torch.cat
    assert ref.requires_grad
    assert_close(blocked, ref.detach(), 2e-2, "blocked-memory inference vs differentiable path")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_layernorm_blocked_output(api, dtype):
    from tvt_b200 import ops
    B, S, d, blocks = 6, 33, 256, 3
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(B * S, d, device=DEV, generator=g).to(dtype)
    gamma, beta = torch.rand(d, device=DEV, generator=g) + 0.5, torch.randn(d, device=DEV, generator=g)
    dense, _, _ = ops.layernorm_fwd(x, gamma, beta)
    buf = torch.full((B, blocks * S, d), 7.0, device=DEV, dtype=dtype)
    view, _, _ = ops.layernorm_fwd(x, gamma, beta, out=(buf, 1))
    assert torch.equal(view.reshape(B * S, d), dense) and torch.equal(buf[:, S:2 * S].reshape(B * S, d), dense)
    assert bool((buf[:, :S] == 7.0).all()) and bool((buf[:, 2 * S:] == 7.0).all())      # the neighbouring blocks are untouched
    with pytest.raises(Exception, match="out buffer"):
        ops.layernorm_fwd(x, gamma, beta, out=(buf[:, :S + 1], 0))
