"""Helpers shared by the parity tests.

Every comparison made through ``assert_close`` / ``grads_close`` with a ``case=`` name is also RECORDED (error,
tolerance, the stock-PyTorch-bf16 yardstick error and whether the yardstick clause was needed), so that the
softening of the bf16 gradient bar is auditable: ``conftest.pytest_terminal_summary`` prints one line per case
(``N params, worst e, K exemptions, worst exempt ratio``) — which therefore lands in the driver's GPUTEST tail —
and writes the full per-parameter table to ``profiles/parity_report.json`` (and ``gpurun_out/`` on a GPU box).
"""
import torch

# recorded comparisons: list of dicts (see _record)
REPORT = []
# free-form one-line notes for the terminal summary (top-1 agreement counts, margin histograms ...)
NOTES = []
# set by conftest's autouse fixture for gpu-marked tests: the default `case` of assert_close / grads_close
CURRENT_CASE = [None]

# The only parameters that may use the yardstick clause, by class (every use is listed in the parity report):
#   "linear1.", "reason.relation."  ReLU-gated layers (encoder FFN-in, the pyramid's relation MLPs): a pre-activation within
#                                   rounding distance of zero lands on the other side of the gate, and one flipped gate is a
#                                   finite error of that layer's weight / bias gradient;
#   "cls"                           per-batch-slot CLS rows: single-token gradients, no averaging over tokens;
#   "expert_encoder."               the input projection of a memory stream (its only gradient path is the cross-attention
#                                   softmax over the other modality: small, cancellation-prone sums).
# Anything else exceeding the tolerance fails outright, whatever stock PyTorch does.
EXEMPTIBLE = ("linear1.", "reason.relation.", "cls", "expert_encoder.")


def _record(case, what, err, tol, yard=None, exempt=False, limit=None):
    case = case if case is not None else CURRENT_CASE[0]
    if case is None:
        return
    REPORT.append({"case": case, "what": what, "err": float(err), "tol": float(tol),
                   "yardstick_err": None if yard is None else float(yard), "exempt": bool(exempt),
                   "limit": float(tol if limit is None else limit)})


class stock_reduced_precision:
    """Context for the YARDSTICK run of the fp32 oracle: what stock PyTorch gives in the reduced-precision mode that
    corresponds to the tvt mode under test — ``torch.autocast(bf16)`` for the bf16 mode, TF32 tensor-core matmuls
    (``allow_tf32``: the library's own "fp32 data, tensor cores, fp32 accumulate" mode) for the fp32-accumulate mode."""

    def __init__(self, precision):
        self.precision = precision

    def __enter__(self):
        if self.precision == "bf16":
            self.ctx = torch.autocast("cuda", dtype=torch.bfloat16)
            self.ctx.__enter__()
        else:
            self.saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        return self

    def __exit__(self, *exc):
        if self.precision == "bf16":
            return self.ctx.__exit__(*exc)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.saved
        return False


YARD_PRECISION = ["bf16"]      # which reduced-precision mode `reduced(fn)` runs under (set by the tests' _yardstick helpers)


def reduced(fn):
    """Run a forward under the stock reduced-precision mode of the case being checked (losses are evaluated outside)."""
    with stock_reduced_precision(YARD_PRECISION[0]):
        return fn()


def rel_err(a, b):
    """Normwise relative error ||a - b|| / ||b|| in float64."""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def assert_close(a, b, tol, what="", case=None):
    e = rel_err(a.cpu(), b.cpu())
    _record(case, what, e, tol)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


def copy_state(dst, src):
    """Load src's state_dict into dst (same key set required) and return dst."""
    missing, unexpected = dst.load_state_dict(src.state_dict(), strict=True)
    assert not missing and not unexpected
    return dst


def grads_close(mod, ref, tol, what="", skip=(), yard=None, slack=1.5, case=None, exemptible=EXEMPTIBLE):
    """Every parameter gradient of `mod` matches `ref`'s within normwise tolerance `tol`.

    `yard` (optional) is a copy of the fp32 oracle whose gradients were computed by stock PyTorch in the matching
    reduced-precision mode (`stock_reduced_precision`: autocast(bf16) for the bf16 mode, TF32 matmuls for the
    fp32-accumulate mode): a parameter whose name contains one of `exemptible` may exceed `tol` only if stock
    PyTorch does so too, and then by at most `slack` x the yardstick's own error (a single flipped ReLU gate and
    single-token CLS rows are inherently noisy below true-fp32 precision for any implementation).  Every comparison is recorded
    under `case` for the parity report.  Returns (worst name, worst error, number of yardstick exemptions)."""
    rp = dict(ref.named_parameters())
    yp = dict(yard.named_parameters()) if yard is not None else {}
    worst, exempt = ("", 0.0), 0
    failures = []
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
        limit, ye, used = tol, None, False
        if name in yp and yp[name].grad is not None:
            ye = rel_err(yp[name].grad.cpu(), g_ref.cpu())
        if e > tol and ye is not None and any(s in name for s in exemptible):
            limit = max(tol, slack * ye)
            used = True
            exempt += 1
        _record(case, "grad " + name, e, tol, ye, used, limit)
        if e > limit:
            failures.append(f"{what}{name}: gradient relative error {e:.3e} > {limit:.1e}"
                            + (f" (stock yardstick {ye:.3e})" if ye is not None else ""))
    assert not failures, "; ".join(failures)
    return worst[0], worst[1], exempt


# module-level tests are summarised on the terminal; the single-kernel tests of tests/test_kernels_gpu.py only in the JSON
KERNEL_CASES = set()


def summarize():
    """Per-case one-liners + the full table (for conftest.pytest_terminal_summary)."""
    cases = {}
    for r in REPORT:
        cases.setdefault(r["case"], []).append(r)
    lines = []
    for case, rows in cases.items():
        if case.split("[")[0] in KERNEL_CASES:
            continue                          # single-kernel checks: in the JSON table only
        grads = [r for r in rows if r["what"].startswith("grad ")]
        other = [r for r in rows if not r["what"].startswith("grad ")]
        ex = [r for r in grads if r["exempt"]]
        parts = [f"parity[{case}]:"]
        if other:
            shown = ", ".join(f"{r['what']} {r['err']:.2e}" for r in other[:6])
            if len(other) > 6:
                w = max(other, key=lambda r: r["err"])
                shown += f", ... {len(other)} checks, worst {w['err']:.2e} ({w['what']})"
            parts.append(shown)
        if grads:
            w = max(grads, key=lambda r: r["err"])
            parts.append(f"{len(grads)} param grads, worst {w['err']:.2e} ({w['what'][5:]}), tol {w['tol']:.0e}, "
                         f"{len(ex)} yardstick exemptions")
            if ex:
                wr = max(ex, key=lambda r: r["err"] / max(r["yardstick_err"], 1e-30))
                parts.append(f"worst exempt ratio {wr['err'] / max(wr['yardstick_err'], 1e-30):.2f}x stock "
                             f"({wr['what'][5:]}: {wr['err']:.2e} vs {wr['yardstick_err']:.2e})")
        lines.append(" ".join(parts))
    return lines
