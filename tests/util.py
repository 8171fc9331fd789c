"""Helpers shared by the parity tests."""
import torch


def rel_err(a, b):
    """Normwise relative error ||a - b|| / ||b|| in float64."""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def assert_close(a, b, tol, what=""):
    e = rel_err(a.cpu(), b.cpu())
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


def copy_state(dst, src):
    """Load src's state_dict into dst (same key set required) and return dst."""
    missing, unexpected = dst.load_state_dict(src.state_dict(), strict=True)
    assert not missing and not unexpected
    return dst


def grads_close(mod, ref, tol, what="", skip=(), yard=None, slack=1.5):
    """Every parameter gradient of `mod` matches `ref`'s within normwise tolerance `tol`.

    `yard` (optional) is a copy of the fp32 oracle whose gradients were computed under stock
    torch.autocast(bf16): a parameter may exceed `tol` only if stock PyTorch bf16 does so too, and then by at
    most `slack` x the yardstick's own error (ReLU-gate flips and single-token CLS rows are inherently noisy
    in bf16 for any implementation).  Returns (worst name, worst error, number of yardstick exemptions)."""
    rp = dict(ref.named_parameters())
    yp = dict(yard.named_parameters()) if yard is not None else {}
    worst, exempt = ("", 0.0), 0
    for name, p in mod.named_parameters():
        if any(s in name for s in skip):
            continue
        g_ref = rp[name].grad
        if g_ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{what}{name}: reference has no gradient"
            continue
        assert p.grad is not None, f"{what}{name}: missing gradient"
        if float(g_ref.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) < 1e-6, f"{what}{name}: expected zero gradient"
            continue
        e = rel_err(p.grad.cpu(), g_ref.cpu())
        if e > worst[1]:
            worst = (name, e)
        limit = tol
        if e > tol and name in yp and yp[name].grad is not None:
            limit = max(tol, slack * rel_err(yp[name].grad.cpu(), g_ref.cpu()))
            exempt += 1
        assert e <= limit, f"{what}{name}: gradient relative error {e:.3e} > {limit:.1e}"
    return worst[0], worst[1], exempt
