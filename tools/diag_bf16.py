"""Diagnostic: per-parameter gradient error of the bf16 path vs the fp32 oracle, next to what stock
torch bf16 autocast gives on the same oracle module (yardstick for inherent bf16 noise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import copy
import torch
import tvt_b200
from tvt_b200 import hostapi
from oracle import param
from util import rel_err, copy_state

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
KIND = sys.argv[3] if len(sys.argv) > 3 else "ptn"
torch.manual_seed(1130)
gen = torch.Generator().manual_seed(1130)
if KIND == "ptn":
    cfg = dict(batch_size=B, seq_len=16, cls=1, dropout=0.0, input_dimension=256, nhead=4, nhid=512, nlayers=L,
               model="ptn", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15)
    ref = param.SimpleTransformer(**cfg).to(DEV)
    mod = copy_state(hostapi.SimpleTransformer(precision="bf16", **cfg), ref).to(DEV)
    ac = copy.deepcopy(ref)
    x = torch.randn(B, 16, 2, 256, generator=gen).to(DEV)
    y = (torch.rand(B, 15, generator=gen) < 0.15).float().to(DEV)
    ref.criterion(ref.ptn(x), y).backward()
    mod.training_step({"experts": x, "label": y}, 0).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la = ac.criterion(ac.ptn(x).float(), y)
    la.backward()
else:
    kw = dict(d=896, nhead=2, nhid=512, nlayers=L, dropout=0.0, seq_len=14, n_classes=19)
    ref = param.FrameStream(**kw).to(DEV)
    mod = copy_state(hostapi.FrameStream(precision="bf16", **kw), ref).to(DEV)
    ac = copy.deepcopy(ref)
    feats = torch.randn(B, 14, 896, generator=gen).to(DEV)
    y = (torch.rand(B, 19, generator=gen) < 0.15).float().to(DEV)
    torch.nn.functional.binary_cross_entropy_with_logits(ref(feats), y).backward()
    mod.training_step((y, feats), 0).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la = torch.nn.functional.binary_cross_entropy_with_logits(ac(feats).float(), y)
    la.backward()
rp, apd = dict(ref.named_parameters()), dict(ac.named_parameters())
print(f"B={B} L={L}  {'param':60s} {'ours':>10s} {'autocast':>10s}")
for n, p in mod.named_parameters():
    if p.grad is None or rp[n].grad is None or "encoder_layers" in n or "mlp_encoder" in n:
        continue
    print(f"{n:66s} {rel_err(p.grad, rp[n].grad):10.2e} {rel_err(apd[n].grad, rp[n].grad):10.2e}")
