"""Regenerate tests/golden/reference_outputs.pt by running the UNMODIFIED reference classes.

Run in the container that has /root/reference:   python -m oracle.make_golden
For every case the reference class and the oracle.param restatement are built from the same torch seed
(so they draw identical initial weights), fed the same seeded input, and required to agree before the
REFERENCE's outputs are frozen.  tests/test_oracle_golden.py then re-checks oracle.param against the
frozen vectors wherever the tests run (the GPU box has no /root/reference).  Test infrastructure only.
"""
import ast
import contextlib
import io
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import param, ref_loader

SEED = 1130  # the reference's own seed, src/main.py:25
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_outputs.pt")


def _quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints shapes inside ptn()
        return fn(*a)


def _grad_summary(model):
    return {n: p.grad.detach().double().norm().float() for n, p in model.named_parameters() if p.grad is not None}


def _frame_transformer_classes():
    """PositionalEncoding and TransformerBase exactly as written in src/models/frame_transformer.py:19-47
    (the rest of that file needs torchvision weights from the network)."""
    import pytorch_lightning as pl
    import math
    path = os.path.join(ref_loader.REF_ROOT, "src/models/frame_transformer.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("PositionalEncoding", "TransformerBase")]
    ns = {"pl": pl, "nn": nn, "torch": torch, "math": math,
          "TransformerEncoder": nn.TransformerEncoder, "TransformerEncoderLayer": nn.TransformerEncoderLayer}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["PositionalEncoding"], ns["TransformerBase"]


def case_ptn(ref):
    cfg = dict(batch_size=2, seq_len=4, cls=1, dropout=0.0, input_dimension=2048, nhead=8, nhid=32, nlayers=1,
               model="ptn", learning_rate=1e-3, momentum=0.0, weight_decay=0.0)
    torch.manual_seed(SEED)
    m_ref = ref.transformer.SimpleTransformer(**cfg)
    torch.manual_seed(SEED)
    m_or = param.SimpleTransformer(**cfg)
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(2, 4, 3, 2048, generator=g)          # (BATCH, SEQ, EXPERTS, DIM); expert 2 has no encoder
    y = (torch.rand(2, 15, generator=g) < 0.15).float()
    out = {}
    for name, m in (("ref", m_ref), ("or", m_or)):
        logits = _quiet(m.ptn, x)
        loss = m.criterion(logits, y)
        loss.backward()
        out[name] = (logits.detach(), loss.detach(), _grad_summary(m))
    assert torch.allclose(out["ref"][0], out["or"][0], atol=1e-6), "oracle ptn != reference ptn"
    assert torch.allclose(out["ref"][1], out["or"][1], atol=1e-7)
    for k, v in out["ref"][2].items():
        assert torch.allclose(v, out["or"][2][k], rtol=1e-5, atol=1e-9), k
    assert set(m_ref.state_dict().keys()) == set(m_or.state_dict().keys())
    return {"cfg": cfg, "logits": out["ref"][0], "loss": out["ref"][1], "grad_norms": out["ref"][2],
            "state_dict_keys": sorted(m_ref.state_dict().keys())}


def case_posenc(ref):
    pe_ref = ref.transformer.PositionalEncoding(64, 0.0, max_len=9).pe
    assert torch.equal(pe_ref, param.PositionalEncoding(64, 0.0, max_len=9).pe)
    return {"pe": pe_ref.clone()}


def case_reasoning(ref):
    torch.manual_seed(SEED)
    r_ref = ref.tpn.Reasoning(num_segments=1, num_frames=9, num_class=15, img_dim=32).eval()
    torch.manual_seed(SEED)
    r_or = param.Reasoning(num_segments=1, num_frames=9, num_class=15, img_dim=32).eval()
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(3, 9, 32, generator=g)
    o_ref, o_or = r_ref(x), r_or(x)
    assert torch.allclose(o_ref, o_or, atol=1e-7)
    groups = {gp: ref.tpn.sum_group(x, gp) for gp in (2, 3, 4)}
    for gp, v in groups.items():
        assert torch.equal(v, param.sum_group(x, gp))
    return {"out": o_ref.detach(), "sum_group": groups}


def case_vit(ref):
    torch.manual_seed(SEED)
    t_ref = ref.vit.Transformer(32, 2, 2, 16, 64).eval()
    torch.manual_seed(SEED)
    t_or = param.VitTransformer(32, 2, 2, 16, 64).eval()
    assert list(t_ref.state_dict().keys()) == list(t_or.state_dict().keys())
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(2, 5, 32, generator=g)
    o_ref, o_or = t_ref(x), t_or(x)
    assert torch.allclose(o_ref, o_or, atol=1e-6)
    return {"out": o_ref.detach()}


def case_spatial_pyramid(ref):
    torch.manual_seed(SEED)
    low, mid, high = ref.tpn.Feature_Pyramid_low(), ref.tpn.Feature_Pyramid_Mid(), ref.tpn.Feature_Pyramid_High()
    sp = param.SpatialPyramid()
    sp.pyramid_low["channels_reduce"].load_state_dict(low.channels_reduce.state_dict())
    sp.pyramid_mid["channels_reduce"].load_state_dict(mid.channels_reduce.state_dict())
    g = torch.Generator().manual_seed(SEED)
    xl, xm, xh = (torch.randn(2, c, s, s, generator=g) for c, s in ((128, 28), (256, 14), (512, 7)))
    ref_out = torch.cat((high(xh).flatten(1), mid(xm).flatten(1), low(xl).flatten(1)), dim=-1)   # TPN.py:55-58
    assert torch.allclose(ref_out, sp(xl, xm, xh), atol=1e-6)
    return {"out": ref_out.detach(), "low_w": low.channels_reduce.weight.detach().clone(), "low_b": low.channels_reduce.bias.detach().clone(),
            "mid_w": mid.channels_reduce.weight.detach().clone(), "mid_b": mid.channels_reduce.bias.detach().clone()}


def case_frame_stream(ref):
    PE, TB = _frame_transformer_classes()
    torch.manual_seed(SEED)
    pe_ref, tb_ref = PE(64, 0.0, max_len=6), TB(64, 128, 2, 32, 2, 0.0)
    head_ref = nn.Sequential(nn.Linear(64, 512), nn.GELU(), nn.Linear(512, 128), nn.GELU(), nn.Linear(128, 19))  # frame_transformer.py:106
    torch.manual_seed(SEED)
    fs = param.FrameStream(d=64, nhead=2, nhid=32, nlayers=2, dropout=0.0, seq_len=6, n_classes=19)
    g = torch.Generator().manual_seed(SEED)
    feats = torch.randn(3, 6, 64, generator=g)
    data = tb_ref(pe_ref(feats.permute(1, 0, 2))).permute(1, 0, 2)       # frame_transformer.py:205-209
    out_ref = head_ref(data[:, 0])                                         # :179
    assert torch.allclose(out_ref, fs(feats), atol=1e-6)
    teacher = torch.randn(3, 19, generator=g)
    target = (torch.rand(3, 19, generator=g) < 0.15).float()
    distil = nn.CrossEntropyLoss()(out_ref, torch.argmax(teacher, dim=-1))  # :250
    base = nn.BCEWithLogitsLoss()(out_ref, target)                          # :251
    cos = nn.CosineSimilarity(dim=1)(out_ref, teacher)[0]                   # :257
    loss, parts = param.distill_loss(fs(feats), teacher, target)
    assert torch.allclose(loss, base + distil, atol=1e-6) and torch.allclose(parts["cos"], cos, atol=1e-6)
    return {"logits": out_ref.detach(), "teacher": teacher, "target": target, "loss": (base + distil).detach(), "cos": cos.detach()}


def case_collab(ref):
    """src/models/collabgating.py on its own nested-list input layout against the stacked-row restatement."""
    torch.manual_seed(SEED)
    c_ref = ref.collab.CollaborativeGating().eval()
    torch.manual_seed(SEED)
    c_or = param.CollaborativeGating().eval()
    for (na, a), (nb, b) in zip(c_ref.state_dict().items(), c_or.state_dict().items()):
        assert na == nb and torch.equal(a, b), (na, nb)
    g = torch.Generator().manual_seed(SEED)
    B, S, dims = 2, 3, (2048, 1024, 128)
    xs = [torch.randn(B, S, D, generator=g) for D in dims]
    nested = [[[x[b, s].reshape(1, -1) for x in xs] for s in range(S)] for b in range(B)]
    with torch.no_grad():
        o_ref, o_or = c_ref(nested), c_or(xs)
    assert o_ref.shape == (B, S, 1024) and torch.allclose(o_ref, o_or, atol=2e-6), float((o_ref - o_or).abs().max())
    return {"xs": xs, "out": o_ref.detach()}


def case_vivit(ref):
    """src/models/vit.py ViViT (small configuration) against the restatement; the reference's output is frozen."""
    kw = dict(image_size=16, patch_size=8, num_classes=5, num_frames=3, dim=32, depth=1, heads=2, dim_head=16)
    torch.manual_seed(SEED)
    v_ref = ref.vit.ViViT(**kw).eval()
    torch.manual_seed(SEED)
    v_or = param.ViViT(**kw).eval()
    assert list(v_ref.state_dict().keys()) == list(v_or.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(v_ref.state_dict().values(), v_or.state_dict().values()))
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(2, 3, 3, 16, 16, generator=g)
    o_ref, o_or = v_ref(x.clone()), v_or(x.clone())
    assert torch.allclose(o_ref, o_or, atol=1e-6)
    return {"kw": kw, "out": o_ref.detach(), "state_dict_keys": list(v_ref.state_dict().keys())}


def case_frame_transformer(ref):
    """The UNMODIFIED FrameTransformer(**config.yaml) in its one runnable mode ("vid"): real R(2+1)D-18 backbone
    (torchvision's architecture, seeded random weights: no network), forward + training loss on a seeded batch, against
    oracle.param.FrameTransformer with the same backbone module and weights.  Also freezes the reference's constructor
    contract: config.yaml's keys, the hyper-parameters it stores and its state_dict key / shape list."""
    import yaml
    ft = ref_loader.load_frame_transformer()
    cfg = yaml.safe_load(open(os.path.join(ref_loader.REF_ROOT, "src/config.yaml")))
    torch.manual_seed(SEED)
    m_ref = ft.FrameTransformer(**cfg)
    g = torch.Generator().manual_seed(SEED)
    B = cfg["batch_size"]
    vid = torch.randn(B, 13, 12, 3, 112, 112, generator=g) * 0.5
    target = (torch.rand(B, 19, generator=g) < 0.15).double()
    for sub in m_ref.modules():
        if isinstance(sub, nn.Dropout):
            sub.p = 0.0                      # position_encoder / encoder dropout 0.5 is hard-coded (:92,99)
        if isinstance(sub, nn.MultiheadAttention):
            sub.dropout = 0.0                # attention-probability dropout is a float attribute, not a module
    m_ref.train()
    loss = m_ref.training_step((target, None, vid), 0)
    loss.backward()
    logits = m_ref(None, vid).detach()
    m_or = param.FrameTransformer(model="vid", batch_size=B, seq_len=13, cls=1, dropout=0.0, vid_model=m_ref.vid_model)
    sd = {k: v for k, v in m_ref.state_dict().items() if not k.startswith("vid_model.")}
    missing = m_or.load_state_dict(sd, strict=False)
    assert all(k.startswith("vid_model.") for k in missing.missing_keys) and not missing.unexpected_keys
    m_or.train()
    assert torch.allclose(m_or(None, vid), logits, atol=1e-5)
    assert torch.allclose(m_or.loss((target, None, vid)), loss.detach().float(), atol=1e-6)
    head_grad = m_ref.img_mlp_head[4].weight.grad.detach().double().norm().float()
    return {"config": cfg, "hparams": {k: v for k, v in m_ref.hparams.items()},
            "state_dict_shapes": {k: tuple(v.shape) for k, v in m_ref.state_dict().items()},
            "logits": logits, "loss": loss.detach().float(), "head_grad_norm": head_grad,
            "simple_transformer_state_dict_shapes": {k: tuple(v.shape) for k, v in
                                                     ref.transformer.SimpleTransformer(**dict(cfg, model="ptn")).state_dict().items()},
            "tpn_state_dict_keys": {name: list(getattr(ref.tpn, name)().state_dict().keys())
                                    for name in ("Feature_Pyramid_low", "Feature_Pyramid_Mid", "Feature_Pyramid_High", "Reasoning")}}


def main():
    if not ref_loader.available():
        sys.exit("reference not available: golden vectors can only be regenerated where /root/reference exists")
    ref = ref_loader.load()
    gold = {"seed": SEED, "torch": torch.__version__}
    for fn in (case_ptn, case_posenc, case_reasoning, case_vit, case_spatial_pyramid, case_frame_stream, case_collab, case_vivit,
               case_frame_transformer):
        gold[fn.__name__[5:]] = fn(ref)
        print("ok", fn.__name__)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
