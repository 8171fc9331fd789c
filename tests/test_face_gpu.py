"""GPU: the reference-named classes added for the Python face against the oracle —
FrameTransformer in every mode, the trainable spatial pyramid (Feature_Pyramid_low / Mid / High, TPN), ViViT — and a
Trainer-style fit loop over the drop-in modules."""
import copy
import os

import pytest
import torch
import torch.nn.functional as F

from util import assert_close, copy_state, grads_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-3, "bf16": 2e-2}
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.pt")


@pytest.fixture(scope="module")
def api():
    import tvt_b200
    from tvt_b200 import hostapi
    assert tvt_b200.capi.load().tvt_device_check() == 0
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return hostapi


def _no_dropout(*mods):
    for m in mods:
        for sub in m.modules():
            if isinstance(sub, torch.nn.Dropout):
                sub.p = 0.0
            if isinstance(sub, torch.nn.MultiheadAttention):
                sub.dropout = 0.0


def _ac(fn):
    import util
    return util.reduced(fn)


def _targets(B, C, gen):
    y = (torch.rand(B, C, generator=gen) < 0.15).float()
    y[torch.arange(B), torch.randint(0, C, (B,), generator=gen)] = 1.0
    return y


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("model", ["vid", "frame", "distil", "sum", "sum_residual", "pre_modal"])
def test_frame_transformer_modes_parity(api, model, precision):
    """FrameTransformer at its real widths (d = 896; video stream 2 heads x 448, image stream 4 heads x 224, ff 512 / 896,
    4 layers, 14 scene tokens) on backbone FEATURES, every mode of frame_transformer.py:136-180, loss of :246-282."""
    from oracle import param
    B = 4 if precision == "fp32" else 96
    torch.manual_seed(1130)
    ref = param.FrameTransformer(model=model, batch_size=B, seq_len=13, cls=1, dropout=0.0).to(DEV)
    mod = api.FrameTransformer(model=model, batch_size=B, seq_len=13, cls=1, vid_model="features", img_model="features",
                               precision=precision, opt="adamW", learning_rate=1e-4, weight_decay=0.0)
    copy_state(mod, ref).to(DEV)
    _no_dropout(ref, mod)
    gen = torch.Generator().manual_seed(1130)
    img = torch.relu(torch.randn(B, 13, 896, generator=gen) * 0.5).to(DEV)
    vid = torch.relu(torch.randn(B, 13, 896, generator=gen) * 0.5).to(DEV)
    y = _targets(B, 19, gen).to(DEV)
    if model == "distil":            # decisive teacher: argmax(teacher) must not hinge on rounding of near-tied logits
        with torch.no_grad():
            for m in (ref, mod):
                m.img_mlp_head[4].bias[5] += 1.0
    ref.train(); mod.train()
    loss_r = ref.loss((y, img, vid))
    loss_r.backward()
    loss = mod.training_step((y.double(), img, vid), 0)          # the loader's float64 targets
    loss.backward()
    out_r, out = ref(img, vid), mod(img, vid)
    if model == "distil":
        assert_close(out[0], out_r[0], TOL[precision], "student logits")
        assert_close(out[1], out_r[1], TOL[precision], "teacher logits")
    else:
        assert_close(out, out_r, TOL[precision], "logits")
    assert_close(loss, loss_r, TOL[precision], "loss")
    import util
    yard = copy.deepcopy(ref)
    yard.zero_grad(set_to_none=True)
    util.YARD_PRECISION[0] = precision

    def run(m):
        if model == "distil":
            s, t = _ac(lambda: m(img, vid))
            return m.criterion(s.float(), y) + m.distil_criterion(s.float(), torch.argmax(t.float(), dim=-1))
        return m.criterion(_ac(lambda: m(img, vid)).float(), y)
    run(yard).backward()
    worst = grads_close(mod, ref, TOL[precision], f"{model} ", yard=yard)
    print("FrameTransformer", model, precision, "worst grad", worst)
    # evaluation hooks feed the callbacks' side channel (frame_transformer.py:331-333,364-366)
    mod.eval()
    with torch.no_grad():
        mod.validation_step((y.double(), img, vid), 0)
        mod.test_step((y.double(), img, vid), 0)
    assert len(mod.running_logits) == 2 and mod.running_logits[0].shape == (B, 19) and mod.running_labels[0].dtype == torch.int32
    assert float(mod.running_logits[0].min()) >= 0.0 and float(mod.running_logits[0].max()) <= 1.0


def test_frame_transformer_from_config_matches_reference_golden(api):
    """``FrameTransformer(**config.yaml)`` — the exact call of src/main.py:44 — with its default (torchvision, no download)
    R(2+1)D-18 backbone, seeded like the reference run that was frozen in tests/golden: same initial weights, and the
    sm_100a path behind the backbone reproduces the UNMODIFIED reference's logits and training loss."""
    gold = torch.load(GOLD, weights_only=False)
    g = gold["frame_transformer"]
    torch.manual_seed(gold["seed"])
    mod = api.FrameTransformer(precision="fp32", **g["config"]).to(DEV)
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == g["state_dict_shapes"]
    _no_dropout(mod)
    mod.train()
    gen = torch.Generator().manual_seed(gold["seed"])
    B = g["config"]["batch_size"]
    vid = (torch.randn(B, 13, 12, 3, 112, 112, generator=gen) * 0.5).to(DEV)
    target = (torch.rand(B, 19, generator=gen) < 0.15).double().to(DEV)
    loss = mod.training_step((target, None, vid), 0)
    loss.backward()
    assert_close(loss, g["loss"], 1e-3, "golden loss")
    got = mod.img_mlp_head[4].weight.grad.double().norm().item()
    assert abs(got - float(g["head_grad_norm"])) <= 2e-3 * float(g["head_grad_norm"])
    assert mod.vid_model.backbone.stem[0].weight.grad is not None          # the backbone trains (VidResNet has no no_grad)
    with torch.no_grad():
        assert_close(mod(None, vid), g["logits"], 1e-3, "golden logits")


@pytest.mark.parametrize("precision,dtype", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
@pytest.mark.parametrize("frames", [20, 7])
def test_spatial_pyramid_trains(api, precision, dtype, frames):
    """Feature_Pyramid_low / Mid / High + concat (TPN.py:2-40,55-58) forward AND backward: gradients of the trainable
    1x1 convs and of the CNN feature maps (TPN's trunk trains).  frames = 7 takes the generic kernel for the 7x7 level
    (rows not a multiple of the 16-byte period), frames = 20 the TMA-bulk tile ring for all three levels."""
    from oracle import param
    torch.manual_seed(1130)
    ref = param.SpatialPyramid().to(DEV)
    mod = copy_state(api.SpatialPyramid(precision=precision), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    maps = [torch.randn(frames, c, s, s, generator=gen).to(DEV) for c, s in ((128, 28), (256, 14), (512, 7))]
    w = torch.randn(frames, 896, generator=gen).to(DEV)
    maps_r = [m.clone().requires_grad_(True) for m in maps]
    maps_m = [m.to(dtype).requires_grad_(True) for m in maps]
    out_r = ref(*maps_r)
    (out_r * w).sum().backward()
    out = mod(*maps_m)
    (out * w).sum().backward()
    assert_close(out, out_r, TOL[precision], "pyramid features")
    for name in ("pyramid_low", "pyramid_mid"):
        conv_r, conv_m = getattr(ref, name)["channels_reduce"], getattr(mod, name).channels_reduce
        assert_close(conv_m.weight.grad, conv_r.weight.grad, TOL[precision], name + " conv weight grad")
        assert_close(conv_m.bias.grad, conv_r.bias.grad, TOL[precision], name + " conv bias grad")
    assert mod.pyramid_high.channels_reduce.weight.grad is None          # the High level never applies its conv (:24-26)
    for a, b, name in zip(maps_m, maps_r, ("low", "mid", "high")):
        assert_close(a.grad.float(), b.grad, TOL[precision], name + " map grad")
    # the reference-named single-level modules keep the reference's output shape
    assert api.Feature_Pyramid_Mid(precision).to(DEV)(maps[1]).shape == (frames, 256, 1, 1)


def test_tpn_forward_with_injected_trunk(api):
    """TPN.forward (TPN.py:52-61): trunk -> pyramid -> Reasoning over the 20 frames of one clip."""
    from oracle import param

    class Trunk(torch.nn.Module):               # stands in for custom_resnet.resnet34: three maps per frame
        def forward(self, x):
            g = torch.Generator().manual_seed(int(x.sum().item()) % 1000)
            return tuple(torch.randn(x.shape[0], c, s, s, generator=g).to(x.device) for c, s in ((128, 28), (256, 14), (512, 7)))

    torch.manual_seed(1130)
    tpn = api.TPN(net=Trunk(), precision="fp32").to(DEV).eval()
    sp_ref, reason_ref = param.SpatialPyramid().to(DEV), param.Reasoning().to(DEV).eval()
    sp_ref.load_state_dict({k: v for k, v in tpn.state_dict().items() if k.startswith("pyramid_")})
    reason_ref.load_state_dict({k[len("reason."):]: v for k, v in tpn.state_dict().items() if k.startswith("reason.")})
    x = torch.ones(20, 3, 8, 8, device=DEV)
    out = tpn(x)
    want = reason_ref(sp_ref(*Trunk()(x)).unsqueeze(0))
    assert out.shape == (1, 15)
    assert_close(out, want, 1e-3, "TPN output")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vivit_parity(api, precision):
    """src/models/vit.py ViViT (:79-128) at a small image size with its default width (dim 192, 3 heads x 64, depth 2)."""
    from oracle import param
    kw = dict(image_size=32, patch_size=8, num_classes=10, num_frames=6, depth=2)
    torch.manual_seed(1130)
    ref = param.ViViT(**kw).to(DEV)
    mod = copy_state(api.vit.ViViT(precision=precision, **kw), ref).to(DEV)
    gen = torch.Generator().manual_seed(1130)
    b = 4 if precision == "fp32" else 32
    x = torch.randn(b, 6, 3, 32, 32, generator=gen).to(DEV)
    w = torch.randn(b, 10, generator=gen).to(DEV)
    (ref(x.clone()) * w).sum().backward()
    out = mod(x)
    (out * w).sum().backward()
    assert_close(out, ref(x.clone()), TOL[precision], "ViViT logits")
    import util
    yard = copy.deepcopy(ref)
    yard.zero_grad(set_to_none=True)
    util.YARD_PRECISION[0] = precision
    (_ac(lambda: yard(x.clone())).float() * w).sum().backward()
    grads_close(mod, ref, TOL[precision], "vivit ", yard=yard)


def test_vivit_matches_reference_golden(api):
    gold = torch.load(GOLD, weights_only=False)
    g = gold["vivit"]
    torch.manual_seed(gold["seed"])
    mod = api.vit.ViViT(precision="fp32", **g["kw"]).to(DEV).eval()          # same RNG stream as the reference
    gen = torch.Generator().manual_seed(gold["seed"])
    x = torch.randn(2, 3, 3, 16, 16, generator=gen).to(DEV)
    assert_close(mod(x), g["out"], 1e-3, "ViViT golden")


def test_trainer_fit_drives_the_drop_in_modules(api):
    """A Trainer-style run (compat.Trainer when pytorch_lightning is absent: same hook names and order) over the drop-in
    classes for two optimisation steps + validation, with a TransformerEval-like callback reading the side channel."""
    from tvt_b200 import compat

    class Eval(compat.Callback):                 # what callbacks.TransformerEval reads (callbacks.py:34-35,61-62)
        def __init__(self):
            self.seen = []

        def on_validation_epoch_end(self, trainer, pl_module):
            labels = torch.cat(pl_module.running_labels).cpu()
            probs = torch.cat([t for t in pl_module.running_logits if t.shape == pl_module.running_logits[0].shape]).cpu()
            self.seen.append((tuple(labels.shape), tuple(probs.shape)))
            pl_module.running_labels, pl_module.running_logits = [], []

    gen = torch.Generator().manual_seed(1130)
    B = 8
    # (1) SimpleTransformer on the MMX temporal loader's dict batches (MMX_Temporal_dl.py:253-256)
    cfg = dict(batch_size=B, seq_len=12, cls=1, dropout=0.5, input_dimension=256, nhead=4, nhid=512, nlayers=2, model="ptn",
               learning_rate=1e-2, momentum=0.9, weight_decay=0.0, n_classes=15)
    torch.manual_seed(1130)
    st = api.SimpleTransformer(**cfg).to(DEV)
    batches = [{"experts": torch.randn(B, 12, 3, 256, generator=gen).to(DEV), "label": _targets(B, 15, gen).double().reshape(B, 1, 15).to(DEV),
                "path": ["clip"] * B} for _ in range(3)]
    before = st.mlp_head[1].weight.detach().clone()
    cb = Eval()
    trainer = compat.Trainer(max_epochs=1, max_steps=2, callbacks=[cb])
    trainer.fit(st, train_dataloaders=batches, val_dataloaders=batches[:2])
    assert trainer.global_step == 2 and not torch.equal(before, st.mlp_head[1].weight)
    assert cb.seen == [((2 * B, 1, 15), (4 * B, 15))]      # sigmoid AND raw logits are appended (transformer.py:155-158)
    assert torch.isfinite(st.logged["train/loss"])
    # (2) FrameTransformer ("distil", features) on the frame loader's (target, img, vid) tuples (MMX_Light_dl.py:286)
    torch.manual_seed(1130)
    ft = api.FrameTransformer(model="distil", batch_size=B, seq_len=13, cls=1, vid_model="features", img_model="features",
                              opt="adamW", learning_rate=1e-3, weight_decay=0.09).to(DEV)
    fb = [(_targets(B, 19, gen).double().to(DEV), torch.randn(B, 13, 896, generator=gen).to(DEV),
           torch.randn(B, 13, 896, generator=gen).to(DEV)) for _ in range(3)]
    before = ft.img_mlp_head[4].weight.detach().clone()
    cb = Eval()
    trainer = compat.Trainer(max_epochs=1, max_steps=2, callbacks=[cb])
    trainer.fit(ft, train_dataloaders=fb, val_dataloaders=fb[:1])
    assert trainer.global_step == 2 and not torch.equal(before, ft.img_mlp_head[4].weight)
    assert cb.seen == [((B, 19), (B, 19))]
    for key in ("train/loss", "train/distilloss", "train/bass_loss", "train/cossim", "val/loss"):
        assert key in ft.logged and torch.isfinite(torch.as_tensor(ft.logged[key])).all(), key
    trainer.test(ft, dataloaders=fb[:1])
    assert len(ft.running_logits) == 1
